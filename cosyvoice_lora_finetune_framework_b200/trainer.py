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
    """Linear warm-up then cosine decay to min_lr (reference train_joint.py:210-219). `step` is the number of optimiser
    steps already applied: LambdaLR with interval 'step' runs the (k+1)-th step with lambda(k), so the very first
    step of a warm-up has lr = 0. The device-side copy of this rule is optim_advance_kernel (csrc/optim.cu)."""
    if step < warmup_steps:
        return step / max(1, warmup_steps)
    progress = (step - warmup_steps) / max(1, total_steps - warmup_steps)
    return max(min_lr / base_lr, 0.5 * (1 + math.cos(progress * 3.14159)))


class _LossHandle:
    """Loss of a step whose device->host copy is in flight."""

    def __init__(self, ring, idx, event):
        self.ring, self.idx, self.event = ring, idx, event

    def value(self):
        self.event.synchronize()
        return float(self.ring[self.idx])


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
        self.step_count = 0         # host mirror of opt_state[0] (exact unless a step was skipped: see poll_overflow)
        self.micro = 0
        self.kernel_launches_per_step = 0
        # The step counter, learning rate and Adam bias corrections live in DEVICE memory and are advanced by a
        # one-thread kernel inside the step (cvflow_optim_advance): a captured CUDA graph of the whole step needs no
        # per-replay host write, so the host can run any number of replays ahead of the GPU without racing it.
        self.opt_state = torch.zeros(2, device=dev, dtype=torch.int32)     # {steps applied, steps skipped}
        self.hyper = torch.zeros(4, device=dev)                           # {lr, 1-b1^t, sqrt(1-b2^t), apply flag}
        self._skipped_seen = 0
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

        self._chunks = []
        if self.world > 1:
            self._setup_grad_chunks()
        if self.world > 1:      # replicas must start identical (DDP broadcasts rank 0's parameters at construction)
            dist.broadcast(self.ne.param_bucket, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                           group=self.pg)
            if self.extra:
                dist.broadcast(self.xparam, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0,
                               group=self.pg)
            self.ne.mark_dirty()
            self.ne.sync_lora()

    def _setup_grad_chunks(self, n_chunks=3):
        """Overlap the gradient exchange with the backward (what DDP's bucketed hooks do for the reference's modules): the
        flat bucket is cut into chunks of whole attention blocks; the estimator backward finalises a chunk as soon as its
        last-visited block is done and records an event, and optimizer_step allreduces each chunk on a side stream behind
        that event. The chunk holding the first blocks (finished last) is kept small so that little is left exposed."""
        ne = self.ne
        r = ne.lora_r
        if r <= 0 or ne.n_lora <= 0 or int(self.cfm.num_streams) > 1:
            return
        per_block = 3 * (r * 256 + 512 * r)
        nb = ne.n_lora // per_block
        if nb * per_block != ne.n_lora or nb < 4:
            return
        lo = sorted({0, max(1, nb // 8), max(2, nb // 2)})[:n_chunks]
        self._comm_stream = torch.cuda.Stream(device=ne.device)
        evs = [torch.cuda.Event() for _ in lo]
        for e in evs:
            e.record()           # materialises the cudaEvent_t handed to the library
        lo_arr = (C.c_int32 * len(lo))(*lo)
        ev_arr = (C.c_void_p * len(lo))(*[e.cuda_event for e in evs])
        N.check(self.L.cvflow_set_grad_chunks(ne.handle, len(lo), lo_arr, ev_arr), "cvflow_set_grad_chunks")
        hi = lo[1:] + [nb]
        self._chunks = [(a * per_block, b * per_block, e) for a, b, e in zip(lo, hi, evs)]

    def _allreduce_grads(self, g):
        if not self._chunks:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
            return
        cur, side = torch.cuda.current_stream(), self._comm_stream
        for a, b, ev in reversed(self._chunks):      # the chunk of the last blocks is complete first
            side.wait_event(ev)
            with torch.cuda.stream(side):
                dist.all_reduce(g[a:b], op=dist.ReduceOp.SUM, group=self.pg)
        cur.wait_stream(side)

    def current_lr(self):
        """Learning rate the NEXT optimiser step will use (LambdaLR value after `step_count` scheduler steps)."""
        if self.total_steps <= 0:
            return self.lr
        return self.lr * lr_lambda(self.step_count, self.warmup_steps, self.total_steps, self.lr, self.min_lr)

    # -- optimiser state (checkpoint / resume, Lightning's ckpt_path semantics) -------------------------
    def state_dict(self):
        st = int(self.opt_state[0].item())
        out = {'step': st, 'skipped': int(self.opt_state[1].item()), 'm': self.m.detach().cpu(), 'v': self.v.detach().cpu(),
               'loss_scale': float(self.ne.loss_scale)}
        if self.extra:
            out['xm'], out['xv'] = self.xm.detach().cpu(), self.xv.detach().cpu()
        return out

    def load_state_dict(self, sd):
        """Restore Adam moments, the step counter (=> warm-up / cosine position and bias corrections) and the loss
        scale. Call after the model's own state_dict was loaded (the LoRA parameters are views of the flat bucket)."""
        with torch.no_grad():
            if sd['m'].numel() != self.m.numel():
                raise ValueError("optimizer state has %d LoRA moments, this model has %d" % (sd['m'].numel(), self.m.numel()))
            self.m.copy_(sd['m'])
            self.v.copy_(sd['v'])
            if self.extra:
                if 'xm' not in sd or sd['xm'].numel() != self.xm.numel():
                    raise ValueError("optimizer state lacks the upstream (encoder-LoRA) moments of this model")
                self.xm.copy_(sd['xm'])
                self.xv.copy_(sd['xv'])
            self.opt_state[0] = int(sd.get('step', 0))
            self.opt_state[1] = int(sd.get('skipped', 0))
        self.step_count = int(sd.get('step', 0))
        self._skipped_seen = int(sd.get('skipped', 0))
        if 'loss_scale' in sd and float(sd['loss_scale']) != float(self.ne.loss_scale):
            self.ne.loss_scale = float(sd['loss_scale'])
            self._graph = None
        self.ne.mark_dirty()
        self.ne.sync_lora()

    def poll_overflow(self):
        """GradScaler bookkeeping of the reference's '16-mixed' run, polled (one host sync) instead of per step: steps
        whose gradient norm was not finite were skipped on the device (no update, no schedule advance). Returns the
        number skipped since the last poll; on any, the static fp16 loss scale is halved (floor 1) and a captured
        step graph is dropped so that the next call re-captures with the new scale."""
        st = self.opt_state.tolist()
        self.step_count = int(st[0])
        new = int(st[1]) - self._skipped_seen
        self._skipped_seen = int(st[1])
        if new > 0:
            self.found_inf.zero_()
            if self.ne.loss_scale > 1.0:
                self.ne.loss_scale = max(1.0, self.ne.loss_scale * 0.5)
                for r in self.ne.replicas:
                    r.loss_scale = self.ne.loss_scale
                self._graph = None
        return new

    def micro_step(self, x1, mask, mu, spks, cond, prompt_lens=None):
        """Forward + backward of one micro-batch; gradients accumulate in the flat bucket."""
        loss, _ = self.cfm.compute_loss(x1, mask, mu, spks, cond=cond, prompt_lens=prompt_lens)
        (loss / self.accumulate).backward()
        self.micro += 1
        return loss

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

    def optimizer_step(self):
        ne = self.ne
        st = E._stream()
        g = ne.grad_bucket
        if self.extra:
            self._gather_extra()
        if self.world > 1:
            self._allreduce_grads(g)
            if self.extra:
                dist.all_reduce(self.xgrad, op=dist.ReduceOp.SUM, group=self.pg)
        N.check(self.L.cvflow_sumsq(g.data_ptr(), ne.n_lora, self.partials.data_ptr(), self.sumsq.data_ptr(), st),
                "cvflow_sumsq")
        if self.extra:      # one global gradient norm over both buckets (clip_grad_norm_ over all trainable parameters)
            N.check(self.L.cvflow_sumsq(self.xgrad.data_ptr(), self.n_extra, self.xpartials.data_ptr(),
                                        self.xsumsq.data_ptr(), st), "cvflow_sumsq")
            self.sumsq.add_(self.xsumsq)
        # step counter / lr / bias corrections advance on the device (skipped when the gradient norm is not finite)
        N.check(self.L.cvflow_optim_advance(self.opt_state.data_ptr(), self.hyper.data_ptr(), self.sumsq.data_ptr(), None,
                                            1.0 / self.world, float(self.lr), int(self.warmup_steps), int(self.total_steps),
                                            float(self.min_lr), self.betas[0], self.betas[1], st), "cvflow_optim_advance")
        self.step_count += 1
        if self.extra:
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
        ne.sync_lora(need_folded=not ne.trains_unfolded())      # lora_dropout > 0: factor images only (the folded W_eff is unused)
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
        the per-step optimiser scalars come from device memory. Returns the (static) loss tensor.

        With accumulate > 1 (the reference's accumulate_grad_batches, config.py) two graphs are captured: the
        micro-step (forward + backward accumulating into the flat bucket, loss / accumulate) replayed per call, and the
        optimiser tail (allreduce, clip + AdamW, refresh) replayed after every `accumulate`-th call."""
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

            acc = self.accumulate

            def micro_body():
                loss, _ = self.cfm.compute_loss(st["x1"], st["mask"], st["mu"], st["spks"], cond=st["cond"])
                (loss / acc if acc > 1 else loss).backward()
                return loss.detach()

            def body():
                loss = micro_body()
                if acc > 1:                      # warm-up of the two-graph form: a full accumulation window
                    for _ in range(acc - 1):
                        micro_body()
                self.optimizer_step()
                return loss

            self.ne.attach_grads()
            # The two warm-up executions (plans, workspace, NCCL channels) must not count as training: parameters, Adam
            # moments, the device-side step counter, the RNG stream and the LoRA-dropout seed are put back afterwards,
            # so that one call == one optimiser step, exactly like the eager path.
            snap = dict(p=self.ne.param_bucket.clone(), m=self.m.clone(), v=self.v.clone(), o=self.opt_state.clone(),
                        h=self.hyper.clone(), f=self.found_inf.clone(), rng=torch.cuda.get_rng_state(self.ne.device),
                        step=self.step_count, drop=self.ne.dropout_seed_snapshot())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up off the capture stream
                for _ in range(2):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.no_grad():
                self.ne.param_bucket.copy_(snap["p"])
                self.m.copy_(snap["m"])
                self.v.copy_(snap["v"])
                self.opt_state.copy_(snap["o"])
                self.hyper.copy_(snap["h"])
                self.found_inf.copy_(snap["f"])
                self.ne.grad_bucket.zero_()
            self.step_count = snap["step"]
            torch.cuda.set_rng_state(snap["rng"], self.ne.device)
            self.ne.dropout_seed_restore(snap["drop"])
            self.ne.mark_dirty()
            self.ne.sync_lora()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # NB: a loss tensor of an earlier EAGER step that is still referenced keeps that step's autograd graph, and
            # with it the parameters' AccumulateGrad nodes (bound to the stream they were created on), alive; the captured
            # backward would then sync the capturing stream with that uncaptured stream and the capture fails with
            # cudaErrorStreamCaptureIsolation. Drop such references (float(loss) / del loss) before the first graphed step.
            if acc == 1:
                with torch.cuda.graph(g):
                    st["loss"] = body()
            else:
                with torch.cuda.graph(g):
                    st["loss"] = micro_body()
                g2 = torch.cuda.CUDAGraph()
                chunks, self._chunks = self._chunks, []      # the chunk events belong to another capture: one plain allreduce
                try:
                    with torch.cuda.graph(g2):
                        self.optimizer_step()
                finally:
                    self._chunks = chunks
                st["opt_graph"] = g2
            self.step_count = snap["step"]         # capture executes nothing
            self.micro = 0
            st["graph"] = g
            self._graph = st
        st = self._graph
        st["x1"].copy_(x1, non_blocking=True)
        st["mask"].copy_(mask, non_blocking=True)
        st["mu"].copy_(mu, non_blocking=True)
        st["spks"].copy_(spks, non_blocking=True)
        st["cond"].copy_(cond, non_blocking=True)
        st["graph"].replay()
        if "opt_graph" in st:
            self.micro += 1
            if self.micro < self.accumulate:
                return st["loss"]
            st["opt_graph"].replay()
            self.micro = 0
        self.step_count += 1
        return st["loss"]

    # -- pipelined host I/O around the graphed step (the end-to-end path of a real input pipeline) ---------
    def stage_inputs(self, x1, mask, mu, spks, cond):
        """Start copying one step's HOST inputs (pinned memory) to the device on a dedicated copy stream, into one of two
        staging slots, and return at once: the copy of step i+1 overlaps the compute of step i (what a DataLoader with
        pin_memory + prefetch does for the reference, train_joint.py:290-298). Consumed, in order, by train_step_staged."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.ne.device)
            self._slots, self._staged, self._slot_next = [None, None], [], 0
        i = self._slot_next
        self._slot_next ^= 1
        src = dict(x1=x1, mask=mask, mu=mu, spks=spks, cond=cond)
        slot = self._slots[i]
        if slot is None or any(tuple(slot[k].shape) != tuple(v.shape) for k, v in src.items()):
            slot = {k: torch.empty(v.shape, device=self.ne.device, dtype=torch.float32) for k, v in src.items()}
            slot["consumed"] = None
            self._slots[i] = slot
        cs = self._copy_stream
        if slot["consumed"] is not None:
            cs.wait_event(slot["consumed"])      # the step that read this slot last has taken its copy
        with torch.cuda.stream(cs):
            for k, v in src.items():
                slot[k].copy_(v, non_blocking=True)
            slot["ready"] = cs.record_event()
        self._staged.append(slot)

    def train_step_staged(self):
        """One graphed optimiser step on the oldest staged inputs. Returns a handle whose .value() is the step's loss,
        read back through pinned memory without stalling the stream (wait for it one step later)."""
        slot = self._staged.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(slot["ready"])
        loss = self.train_step_graphed(slot["x1"], slot["mask"], slot["mu"], slot["spks"], slot["cond"])
        slot["consumed"] = cur.record_event()
        if not hasattr(self, "_loss_ring"):
            self._loss_ring, self._loss_next = torch.zeros(4).pin_memory(), 0
        j = self._loss_next
        self._loss_next = (j + 1) % 4
        self._loss_ring[j:j + 1].copy_(loss.reshape(1), non_blocking=True)
        return _LossHandle(self._loss_ring, j, cur.record_event())

    def grad_norm(self):
        return float(self.sumsq.sqrt().item()) / self.world
