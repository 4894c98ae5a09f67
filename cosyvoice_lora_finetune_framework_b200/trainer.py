"""Data-parallel flow-LoRA training step (the loop the reference delegates to PyTorch-Lightning:
train_joint.py:105-226,349-360 -- AdamW, linear warm-up -> cosine decay, gradient clip 1.0, gradient
accumulation), re-expressed for one process per GPU:

  per rank   CFM step on its own utterance shard (padded to the shard's own max length, exactly
             what N independent reference data loaders would do -- SURVEY.md section 8e)
  exchange   ONE NCCL allreduce of the flat fp32 LoRA-gradient bucket (4.7 MB at r=8) over
             NVLink/NVSwitch; nothing else crosses ranks (frozen weights are replicated)
  tail       fused global-norm clip + AdamW over the flat bucket, then the W_eff refresh kernel

LoRA parameters UPSTREAM of the estimator (Conformer encoder linear_q/k/v, w_1/w_2 -- the rest of the
reference's flow_lora target list, config.py:207-216) can be handed in as `extra_params`: they receive
their gradients through dL/dmu / dL/dspks of the estimator backward, live in a second flat bucket and
go through the same allreduce, the same global gradient norm and the same fused AdamW kernel.

N-rank result == the average of N single-process reference runs, one per shard.
"""
import ctypes as C
import math

import torch
import torch.distributed as dist

from . import _estimator as E
from . import _native as N


def lr_lambda(step, warmup_steps, total_steps, base_lr, min_lr):
    """Linear warm-up then cosine decay to min_lr (reference train_joint.py:210-219)."""
    if step < warmup_steps:
        return step / max(1, warmup_steps)
    progress = (step - warmup_steps) / max(1, total_steps - warmup_steps)
    return max(min_lr / base_lr, 0.5 * (1 + math.cos(progress * 3.14159)))


class FlowLoRATrainer:
    def __init__(self, cfm, lr=1e-4, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=1.0,
                 warmup_steps=0, total_steps=0, min_lr=1e-6, accumulate=1, process_group=None, extra_params=None):
        self.cfm = cfm
        self.ne = E.native_of(cfm.estimator)
        self.L = E._lib()
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.max_grad_norm = max_grad_norm
        self.warmup_steps, self.total_steps, self.min_lr = warmup_steps, total_steps, min_lr
        self.accumulate = max(1, accumulate)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        dev = self.ne.device
        n = self.ne.n_lora
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.sumsq = torch.zeros(1, device=dev)
        self.partials = torch.zeros(296, device=dev)
        self.found_inf = torch.zeros(1, device=dev, dtype=torch.int32)
        self.step_count = 0
        self.micro = 0
        self.kernel_launches_per_step = 0
        # optimiser scalars that change every step live in device memory so that a captured CUDA graph
        # of the whole step stays valid: the host refreshes them (pinned -> device) before each replay
        self.hyper = torch.zeros(4, device=dev)
        self.hyper_host = torch.zeros(4).pin_memory()
        self._graph = None
        # second flat bucket: trainable parameters upstream of the estimator (become views into it)
        self.extra = [p for p in (extra_params or []) if p.requires_grad]
        self.n_extra = sum(p.numel() for p in self.extra)
        if self.extra:
            with torch.inference_mode(False), torch.no_grad():
                self.xparam = torch.empty(self.n_extra, device=dev, dtype=torch.float32)
                self.xgrad = torch.zeros(self.n_extra, device=dev, dtype=torch.float32)
                self.xviews, off = [], 0
                for p in self.extra:
                    if p.device != dev or p.dtype != torch.float32:
                        raise ValueError("extra_params must be fp32 parameters on %s" % dev)
                    n = p.numel()
                    view = self.xparam[off:off + n].view_as(p)
                    view.copy_(p.data)
                    p.data = view
                    p.grad = self.xgrad[off:off + n].view_as(p)
                    self.xviews.append((p, off, n))
                    off += n
            self.xm = torch.zeros(self.n_extra, device=dev)
            self.xv = torch.zeros(self.n_extra, device=dev)
            self.xsumsq = torch.zeros(1, device=dev)
            self.xpartials = torch.zeros(296, device=dev)

    def current_lr(self):
        if self.total_steps <= 0:
            return self.lr
        return self.lr * lr_lambda(self.step_count, self.warmup_steps, self.total_steps, self.lr, self.min_lr)

    def micro_step(self, x1, mask, mu, spks, cond, prompt_lens=None):
        """Forward + backward of one micro-batch; gradients accumulate in the flat bucket."""
        loss, _ = self.cfm.compute_loss(x1, mask, mu, spks, cond=cond, prompt_lens=prompt_lens)
        (loss / self.accumulate).backward()
        self.micro += 1
        return loss

    def _advance_hyper(self):
        """step counter, learning rate and Adam bias corrections -> device (async, pinned source)."""
        self.step_count += 1
        t = self.step_count
        self.hyper_host[0] = self.current_lr()
        self.hyper_host[1] = 1.0 - self.betas[0] ** t
        self.hyper_host[2] = math.sqrt(1.0 - self.betas[1] ** t)
        self.hyper.copy_(self.hyper_host, non_blocking=True)

    def _gather_extra(self):
        """Upstream gradients back into the flat bucket when autograd (or zero_grad(set_to_none=True)) replaced the
        .grad views."""
        with torch.no_grad():
            for p, off, n in self.xviews:
                view = self.xgrad[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = view
                elif p.grad.data_ptr() != view.data_ptr():
                    view.add_(p.grad)
                    p.grad = view

    def optimizer_step(self, from_graph=False):
        ne = self.ne
        st = E._stream()
        g = ne.grad_bucket
        if self.extra:
            self._gather_extra()
        if self.world > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
            if self.extra:
                dist.all_reduce(self.xgrad, op=dist.ReduceOp.SUM, group=self.pg)
        if not from_graph:
            self._advance_hyper()
        N.check(self.L.cvflow_sumsq(g.data_ptr(), ne.n_lora, self.partials.data_ptr(), self.sumsq.data_ptr(), st),
                "cvflow_sumsq")
        if self.extra:      # one global gradient norm over both buckets (clip_grad_norm_ over all trainable parameters)
            N.check(self.L.cvflow_sumsq(self.xgrad.data_ptr(), self.n_extra, self.xpartials.data_ptr(),
                                        self.xsumsq.data_ptr(), st), "cvflow_sumsq")
            self.sumsq.add_(self.xsumsq)
            N.check(self.L.cvflow_adamw_step(self.xparam.data_ptr(), self.xgrad.data_ptr(), self.xm.data_ptr(),
                                             self.xv.data_ptr(), self.n_extra, self.sumsq.data_ptr(), 1.0 / self.world,
                                             float(self.max_grad_norm), float(self.current_lr()), self.betas[0],
                                             self.betas[1], self.eps, self.wd, max(1, self.step_count),
                                             self.found_inf.data_ptr(), C.c_void_p(self.hyper.data_ptr()), st),
                    "cvflow_adamw_step")
            self.xgrad.zero_()
        N.check(self.L.cvflow_adamw_step(ne.param_bucket.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                         ne.n_lora, self.sumsq.data_ptr(), 1.0 / self.world, float(self.max_grad_norm),
                                         float(self.current_lr()), self.betas[0], self.betas[1], self.eps, self.wd,
                                         max(1, self.step_count), self.found_inf.data_ptr(),
                                         C.c_void_p(self.hyper.data_ptr()), st), "cvflow_adamw_step")
        g.zero_()
        ne.mark_dirty()
        ne.sync_lora()
        self.micro = 0

    def train_step(self, x1, mask, mu, spks, cond, prompt_lens=None):
        loss = self.micro_step(x1, mask, mu, spks, cond, prompt_lens)
        if self.micro >= self.accumulate:
            self.optimizer_step()
        return loss

    # -- whole-step CUDA graph ------------------------------------------------------------------------
    def train_step_graphed(self, x1, mask, mu, spks, cond):
        """One full optimiser step (CFM prep, estimator fwd, loss, bwd, allreduce, clip+AdamW, W_eff
        refresh: ~1,340 kernel launches) replayed as ONE CUDA graph. Shapes must stay fixed; the
        inputs are copied into static buffers, the RNG advances per replay (graph-safe Philox), and
        the per-step optimiser scalars come from device memory. Returns the (static) loss tensor."""
        if self.accumulate != 1:
            raise ValueError("train_step_graphed captures a whole optimiser step: accumulate must be 1")
        if self.extra:
            raise ValueError("train_step_graphed captures the estimator-only step (prepared mu / spks); with "
                             "extra_params use micro_step / optimizer_step around the model's own forward")
        key = (tuple(x1.shape), tuple(spks.shape))
        if self._graph is None or self._graph["key"] != key:
            dev = self.ne.device
            mk = lambda t: torch.empty(t.shape, device=dev, dtype=torch.float32)
            st = dict(key=key, x1=mk(x1), mask=mk(mask), mu=mk(mu), spks=mk(spks), cond=mk(cond))
            for k in ("x1", "mask", "mu", "spks", "cond"):
                st[k].copy_(locals()[k])

            def body():
                loss, _ = self.cfm.compute_loss(st["x1"], st["mask"], st["mu"], st["spks"], cond=st["cond"])
                loss.backward()
                self.optimizer_step(from_graph=True)
                return loss.detach()

            self.ne.attach_grads()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up off the capture stream (plans, workspace, NCCL)
                for _ in range(2):
                    self._advance_hyper()
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["loss"] = body()
            st["graph"] = g
            self._graph = st
        st = self._graph
        st["x1"].copy_(x1, non_blocking=True)
        st["mask"].copy_(mask, non_blocking=True)
        st["mu"].copy_(mu, non_blocking=True)
        st["spks"].copy_(spks, non_blocking=True)
        st["cond"].copy_(cond, non_blocking=True)
        self._advance_hyper()
        st["graph"].replay()
        return st["loss"]

    def grad_norm(self):
        return float(self.sumsq.sqrt().item()) / self.world
