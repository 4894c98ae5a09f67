"""Host glue between the nn.Module tree (modules.py) and the cvflow C ABI (include/cvflow.h).

Everything here is plumbing: one-off weight re-layout with torch ops at bind time, workspace
allocation from the caching allocator, raw pointers + the current stream handed to the library,
and the autograd hooks. The arithmetic of the path lives in csrc/*.cu.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _native as N

_F32 = 2
LOSS_SCALE_DEFAULT = 4096.0   # static loss scale for fp16 gradients (bf16 runs use 1.0)


class EstimatorIO(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_nb", C.c_int32), ("mask", C.c_void_p), ("mask_nb", C.c_int32),
                ("mu", C.c_void_p), ("mu_nb", C.c_int32), ("t", C.c_void_p), ("t_nb", C.c_int32),
                ("spks", C.c_void_p), ("spks_nb", C.c_int32), ("cond", C.c_void_p), ("cond_nb", C.c_int32),
                ("keep", C.c_void_p), ("out", C.c_void_p), ("B", C.c_int32), ("T", C.c_int32),
                ("iso_len", C.c_int32), ("training", C.c_int32)]


class InputGrads(C.Structure):
    _fields_ = [("dx", C.c_void_p), ("dmu", C.c_void_p), ("dspks", C.c_void_p), ("dcond", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("n_mid", C.c_int32), ("dtype", C.c_int32), ("gelu_erf", C.c_int32),
                ("lora_r", C.c_int32), ("lora_scaling", C.c_float)]


_protos_done = False


def _lib():
    global _protos_done
    L = N.lib()
    if not _protos_done:
        vp, i32, i64, f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
        L.cvflow_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.cvflow_destroy.argtypes = [vp]
        L.cvflow_destroy.restype = None
        L.cvflow_bind.argtypes = [vp, C.c_char_p, vp, i64, i32]
        L.cvflow_workspace_bytes.argtypes = [vp, i32, i32, i32]
        L.cvflow_workspace_bytes.restype = i64
        L.cvflow_set_workspace.argtypes = [vp, vp, i64]
        L.cvflow_lora_refresh.argtypes = [vp, vp]
        L.cvflow_lora_refresh_factors.argtypes = [vp, vp]
        L.cvflow_estimator_forward.argtypes = [vp, C.POINTER(EstimatorIO), vp]
        L.cvflow_estimator_backward.argtypes = [vp, vp, f, vp, vp]
        L.cvflow_estimator_backward_inputs.argtypes = [vp, vp, f, vp, C.POINTER(InputGrads), vp]
        L.cvflow_set_lora_dropout.argtypes = [vp, f, C.c_uint64, vp, i64]
        L.cvflow_lora_dropout_seed.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.cvflow_optim_advance.argtypes = [vp, vp, vp, vp, f, f, i32, i32, f, f, f, vp]
        L.cvflow_set_grad_chunks.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(vp)]
        L.cvflow_solve_capture.argtypes = [vp, i32, i32, f, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.cvflow_solve_replay.argtypes = [vp, i32, i32, vp]
        L.cvflow_solve_release.argtypes = [vp]
        L.cvflow_time_embed.argtypes = [vp, vp, i32, vp, vp, i32, vp]
        L.cvflow_launch_count.argtypes = [vp]
        L.cvflow_launch_count.restype = i64
        L.cvflow_cfm_prep.argtypes = [vp, vp, vp, vp, i32, i32, f, vp]
        L.cvflow_cfm_loss.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, f, f, i32, vp, vp]
        L.cvflow_lora_prepare.argtypes = [vp, vp]
        L.cvflow_euler_update.argtypes = [vp, vp, vp, i32, f, i64, vp]
        L.cvflow_sumsq.argtypes = [vp, i64, vp, vp, vp]
        L.cvflow_adamw_step.argtypes = [vp, vp, vp, vp, i64, vp, f, f, f, f, f, f, f, i32, vp, vp, vp]
        L.cvflow_mlp_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp]
        L.cvflow_mlp_backward.argtypes = [vp, vp, vp, vp, vp, i64, i32, i32, vp]
        L.cvflow_attention_scratch_ints.argtypes = [i32, i32]
        L.cvflow_attention_scratch_ints.restype = i64
        L.cvflow_attention_forward.argtypes = [vp, i64, i32, i32, i32, vp, vp, i32, vp, vp, vp]
        L.cvflow_attention_backward.argtypes = [vp, i64, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
        _protos_done = True
    return L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _lin(m):
    """(weight, bias, lora_module_or_None) of an nn.Linear or a LoRALinear wrapper."""
    from .lora import LoRAConv1d, LoRALinear
    if isinstance(m, LoRALinear):
        return m.original_layer.weight, m.original_layer.bias, m
    if isinstance(m, LoRAConv1d):
        raise NotImplementedError("LoRA on 1x1 convs is not part of the fused estimator path")
    return m.weight, m.bias, None


class NativeEstimator:
    """Owns the cvflow handle, the 16-bit weight images, the flat LoRA param/grad buckets and
    the workspace for one ConditionalDecoder."""

    def __init__(self, module, dtype=torch.float16):
        if not torch.cuda.is_available():
            raise RuntimeError("the cvflow estimator needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.L = _lib()
        self.module = module
        self.dtype = dtype
        self.device = next(module.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("ConditionalDecoder parameters must live on a CUDA device (got %s)" % self.device)
        self.keep = []          # tensors the library holds raw pointers to
        self.tensors = {}       # name -> bound tensor (shared with replicas)
        self.grad_slots = {}    # '<...>.grad' name -> (offset, numel, shape) in the flat bucket
        self.replicas = []
        self.handle = C.c_void_p()
        self.ws = None
        self.ws_key = None
        self.loss_scale = LOSS_SCALE_DEFAULT if dtype == torch.float16 else 1.0
        self._dirty = False
        self._merged_version = -1
        with torch.inference_mode(False):   # buckets must be normal (version-tracked) tensors
            self._build()

    def __del__(self):
        try:
            if self.handle:
                self.L.cvflow_destroy(self.handle)
        except Exception:
            pass

    # -- weight binding ------------------------------------------------------------------------
    def _bind(self, name, t):
        assert t.is_cuda and t.is_contiguous(), name
        code = _F32 if t.dtype == torch.float32 else N.dtype_code(t.dtype)
        self.keep.append(t)
        self.tensors[name] = t
        N.check(self.L.cvflow_bind(self.handle, name.encode(), C.c_void_p(t.data_ptr()), t.numel(), code), "cvflow_bind")

    def _h(self, t):
        return t.detach().to(self.dtype).contiguous()

    def _f(self, t):
        return t.detach().float().contiguous()

    def _build(self):
        m = self.module
        n_blocks = len(m.down_blocks[0][1])
        n_mid = len(m.mid_blocks)
        if len(m.down_blocks) != 2 or len(m.up_blocks) != 2:
            raise NotImplementedError("cvflow estimator is built for channels=(256, 256) (two resolutions)")
        stages = ([("down_blocks.%d" % i, m.down_blocks[i]) for i in range(2)] +
                  [("mid_blocks.%d" % i, m.mid_blocks[i]) for i in range(n_mid)] +
                  [("up_blocks.%d" % i, m.up_blocks[i]) for i in range(2)])
        # LoRA discovery
        loras = []
        for _, st in stages:
            for tb in st[1]:
                for pn in ("to_q", "to_k", "to_v"):
                    _, _, lm = _lin(getattr(tb.attn1, pn))
                    if lm is not None:
                        loras.append(lm)
                for other in (tb.attn1.to_out[0], tb.ff.net[0].proj, tb.ff.net[2]):
                    if not isinstance(other, nn.Linear):
                        raise NotImplementedError("the fused estimator supports LoRA on attn1.to_q/to_k/to_v only")
        r = loras[0].r if loras else 0
        scaling = loras[0].scaling if loras else 1.0
        drop_ps = {float(lm.lora_dropout.p) if isinstance(lm.lora_dropout, nn.Dropout) else 0.0 for lm in loras}
        for lm in loras:
            if lm.r != r or lm.scaling != scaling:
                raise NotImplementedError("all LoRA layers must share rank and alpha")
        if len(drop_ps) > 1:
            raise NotImplementedError("all LoRA layers of the estimator must share lora_dropout")
        self.lora_dropout_p = drop_ps.pop() if drop_ps else 0.0
        self._drop_active = 0.0          # dropout rate currently set on the handle (0 when the module is in eval())
        self._drop_mask = None           # explicit keep masks for parity tests (set_debug_dropout_mask)
        self.lora_modules = loras
        self.lora_r = r
        gelu = m.down_blocks[0][1][0].ff.net[0].approximate
        cfg = Config(n_blocks, n_mid, N.dtype_code(self.dtype), 0 if gelu == "tanh" else 1, r, float(scaling))
        self.cfg = cfg
        N.check(self.L.cvflow_create(C.byref(cfg), C.byref(self.handle)), "cvflow_create")

        def conv3_w(w):      # [Cout][Cin][3] -> [Cout][tap*Cin + c]
            return w.permute(0, 2, 1).reshape(w.shape[0], -1)

        def conv3_wd(w):     # dgrad operand: [Cin][tap'*Cout + co] = w[co][ci][2 - tap']
            return w.flip(2).permute(1, 2, 0).reshape(w.shape[1], -1)

        with torch.no_grad():
            self._bind("time.w1", self._f(m.time_mlp.linear_1.weight))
            self._bind("time.b1", self._f(m.time_mlp.linear_1.bias))
            self._bind("time.w2", self._f(m.time_mlp.linear_2.weight))
            self._bind("time.b2", self._f(m.time_mlp.linear_2.bias))
            self._bind("time.proj_w", self._f(torch.cat([st[0].mlp[1].weight for _, st in stages], 0)))
            self._bind("time.proj_b", self._f(torch.cat([st[0].mlp[1].bias for _, st in stages], 0)))
            # flat LoRA parameter / gradient buckets (params become views into the bucket)
            n_lora = sum(lm.lora_A.numel() + lm.lora_B.numel() for lm in loras)
            self.param_bucket = torch.empty(max(n_lora, 1), device=self.device, dtype=torch.float32)
            self.grad_bucket = torch.zeros(max(n_lora, 1), device=self.device, dtype=torch.float32)
            self.n_lora = n_lora
            off = 0
            self.lora_views = []
            for lm in loras:
                for p in (lm.lora_A, lm.lora_B):
                    n = p.numel()
                    view = self.param_bucket[off:off + n].view_as(p)
                    view.copy_(p.data.float())
                    p.data = view
                    self.lora_views.append((p, off, n))
                    off += n
            for S, st in stages:
                res = st[0]
                P = S + ".0"
                w1 = res.block1.block[0].weight
                self._bind(P + ".block1.w", self._h(conv3_w(w1)))
                self._bind(P + ".block1.b", self._f(res.block1.block[0].bias))
                self._bind(P + ".gn1.w", self._f(res.block1.block[1].weight))
                self._bind(P + ".gn1.b", self._f(res.block1.block[1].bias))
                w2 = res.block2.block[0].weight
                self._bind(P + ".block2.w", self._h(conv3_w(w2)))
                self._bind(P + ".block2.b", self._f(res.block2.block[0].bias))
                self._bind(P + ".gn2.w", self._f(res.block2.block[1].weight))
                self._bind(P + ".gn2.b", self._f(res.block2.block[1].bias))
                wr = res.res_conv.weight[:, :, 0]
                self._bind(P + ".res.w", self._h(wr))
                self._bind(P + ".res.b", self._f(res.res_conv.bias))
                self._bind(P + ".block2.wd", self._h(conv3_wd(w2)))
                self._bind(P + ".in.wd", self._h(torch.cat([conv3_wd(w1), wr.t()], 1)))
                for j, tb in enumerate(st[1]):
                    Q = "%s.1.%d" % (S, j)
                    self._bind(Q + ".norm1.w", self._f(tb.norm1.weight))
                    self._bind(Q + ".norm1.b", self._f(tb.norm1.bias))
                    self._bind(Q + ".norm3.w", self._f(tb.norm3.weight))
                    self._bind(Q + ".norm3.b", self._f(tb.norm3.bias))
                    for pn, short in (("to_q", "q"), ("to_k", "k"), ("to_v", "v")):
                        w, b, lm = _lin(getattr(tb.attn1, pn))
                        if b is not None:
                            raise NotImplementedError("attn1 q/k/v with bias is not supported")
                        self._bind(Q + ".w" + short, self._f(w))
                        if lm is not None:
                            self._bind(Q + ".lora_%s.A" % short, lm.lora_A.data)
                            self._bind(Q + ".lora_%s.B" % short, lm.lora_B.data)
                            for pname, p in ((".A.grad", lm.lora_A), (".B.grad", lm.lora_B)):
                                gname = Q + ".lora_%s" % short + pname
                                self._bind(gname, self._grad_view(p))
                                self.grad_slots[gname] = self._grad_slot(p)
                    if r > 0 and all(_lin(getattr(tb.attn1, pn))[2] is not None for pn in ("to_q", "to_k", "to_v")):
                        # [W_eff ; A_cat] and [W_eff^T ; B_blk]: the LoRA factor images ride along as 64 extra
                        # operand rows, so u = x A_cat^T and v = dY B_blk^T fall out of the q/k/v GEMMs
                        wext = torch.zeros(1600, 256, device=self.device, dtype=self.dtype)
                        wtext = torch.zeros(320, 1536, device=self.device, dtype=self.dtype)
                        self._bind(Q + ".weff_ext", wext)
                        self._bind(Q + ".weff_t_ext", wtext)
                        self._bind(Q + ".weff", wext[:1536])
                        self._bind(Q + ".acat16", wext[1536:])
                        self._bind(Q + ".weff_t", wtext[:256])
                        self._bind(Q + ".bblk16", wtext[256:])
                    elif r > 0:
                        raise NotImplementedError("LoRA must wrap to_q, to_k and to_v of every attention block")
                    else:
                        self._bind(Q + ".weff", torch.empty(1536, 256, device=self.device, dtype=self.dtype))
                        self._bind(Q + ".weff_t", torch.empty(256, 1536, device=self.device, dtype=self.dtype))
                    wo, bo, _ = _lin(tb.attn1.to_out[0])
                    self._bind(Q + ".wo", self._h(wo))
                    self._bind(Q + ".wo_t", self._h(wo.t()))
                    self._bind(Q + ".bo", self._f(bo))
                    w1_, b1_, _ = _lin(tb.ff.net[0].proj)
                    self._bind(Q + ".w1", self._h(w1_))
                    self._bind(Q + ".w1_t", self._h(w1_.t()))
                    self._bind(Q + ".b1", self._f(b1_))
                    w2_, b2_, _ = _lin(tb.ff.net[2])
                    self._bind(Q + ".w2", self._h(w2_))
                    self._bind(Q + ".w2_t", self._h(w2_.t()))
                    self._bind(Q + ".b2", self._f(b2_))
            if r > 0 and self.lora_dropout_p > 0:
                # lora_dropout > 0: the branch cannot be folded; per block [W0 | s B_cat] ([1536][320]) for the forward
                # and [W0^T ; B_blk] ([320][1536]) for the dgrad; the LoRA parts are rewritten by the refresh kernel after
                # every optimiser step
                tbs = [(("%s.1.%d" % (S, j)), tb) for S, st in stages for j, tb in enumerate(st[1])]
                nb = len(tbs)
                self.w0d = torch.zeros(nb, 1536, 320, device=self.device, dtype=self.dtype)
                self.w0t_ext = torch.zeros(nb, 320, 1536, device=self.device, dtype=self.dtype)
                for i, (Q, tb) in enumerate(tbs):
                    w0 = torch.cat([_lin(getattr(tb.attn1, pn))[0].detach().float() for pn in ("to_q", "to_k", "to_v")], 0)
                    self.w0d[i, :, :256] = w0.to(self.dtype)
                    self.w0t_ext[i, :256] = w0.t().to(self.dtype)
                    self._bind(Q + ".w0d", self.w0d[i])
                    self._bind(Q + ".w0t_ext", self.w0t_ext[i])
            # resolution changes
            wds = m.down_blocks[0][2].conv.weight
            self._bind("down_blocks.0.2.w", self._h(conv3_w(wds)))
            self._bind("down_blocks.0.2.b", self._f(m.down_blocks[0][2].conv.bias))
            self._bind("down_blocks.0.2.wd_even", self._h(wds[:, :, 1].t()))
            self._bind("down_blocks.0.2.wd_odd", self._h(torch.cat([wds[:, :, 0].t(), wds[:, :, 2].t()], 1)))
            wc = m.down_blocks[1][2].weight
            self._bind("down_blocks.1.2.w", self._h(conv3_w(wc)))
            self._bind("down_blocks.1.2.b", self._f(m.down_blocks[1][2].bias))
            self._bind("down_blocks.1.2.wd", self._h(conv3_wd(wc)))
            wu = m.up_blocks[0][2].conv.weight      # ConvTranspose1d: [Cin][Cout][4]
            self._bind("up_blocks.0.2.w_even", self._h(torch.cat([wu[:, :, 1].t(), wu[:, :, 3].t()], 1)))
            self._bind("up_blocks.0.2.w_odd", self._h(torch.cat([wu[:, :, 0].t(), wu[:, :, 2].t()], 1)))
            self._bind("up_blocks.0.2.b", self._f(m.up_blocks[0][2].conv.bias))
            self._bind("up_blocks.0.2.wd", self._h(wu.permute(0, 2, 1).reshape(256, 1024)))
            wc = m.up_blocks[1][2].weight
            self._bind("up_blocks.1.2.w", self._h(conv3_w(wc)))
            self._bind("up_blocks.1.2.b", self._f(m.up_blocks[1][2].bias))
            self._bind("up_blocks.1.2.wd", self._h(conv3_wd(wc)))
            wf = m.final_block.block[0].weight
            self._bind("final_block.w", self._h(conv3_w(wf)))
            self._bind("final_block.b", self._f(m.final_block.block[0].bias))
            self._bind("final_block.gn.w", self._f(m.final_block.block[1].weight))
            self._bind("final_block.gn.b", self._f(m.final_block.block[1].bias))
            self._bind("final_block.wd", self._h(conv3_wd(wf)))
            wp = m.final_proj.weight[:, :, 0]       # [80][256]
            wp_pad = torch.zeros(128, 256, device=self.device)
            wp_pad[:80] = wp
            self._bind("final_proj.w", self._h(wp_pad))
            self._bind("final_proj.b", self._f(m.final_proj.bias))
            self._bind("final_proj.wt", self._h(wp_pad.t()))
        self.refresh_lora()

    def _grad_view(self, p):
        for q, off, n in self.lora_views:
            if q is p:
                return self.grad_bucket[off:off + n].view_as(p)
        raise KeyError

    def _grad_slot(self, p):
        for q, off, n in self.lora_views:
            if q is p:
                return off, n, tuple(p.shape)
        raise KeyError

    def make_replica(self):
        """A second handle over the SAME weight images (frozen 16-bit operands, merged W_eff, LoRA
        masters) with its own launch plans, workspace, dL/dpred buffer and gradient bucket, so that
        shards of one batch can run concurrently on several CUDA streams."""
        r = NativeEstimator.__new__(NativeEstimator)
        r.L, r.module, r.dtype, r.device = self.L, self.module, self.dtype, self.device
        r.keep, r.tensors, r.grad_slots, r.replicas = [], {}, self.grad_slots, []
        r.handle = C.c_void_p()
        r.ws, r.ws_key = None, None
        r.loss_scale, r._dirty, r._merged_version = self.loss_scale, False, -1
        r.cfg, r.lora_modules, r.lora_r, r.lora_views, r.n_lora = self.cfg, self.lora_modules, self.lora_r, [], self.n_lora
        r.param_bucket = self.param_bucket
        r.grad_bucket = torch.zeros_like(self.grad_bucket)
        N.check(self.L.cvflow_create(C.byref(self.cfg), C.byref(r.handle)), "cvflow_create")
        for name, t in self.tensors.items():
            if name in self.grad_slots:
                off, n, shape = self.grad_slots[name]
                t = r.grad_bucket[off:off + n].view(shape)
            r._bind(name, t)
        N.check(self.L.cvflow_lora_prepare(r.handle, _stream()), "cvflow_lora_prepare")
        r.is_replica = True
        return r

    def shard_handles(self, n):
        """[self, replica_1, ...] of length n (replicas are created once and cached)."""
        while len(self.replicas) < n - 1:
            self.replicas.append(self.make_replica())
        return [self] + self.replicas[: n - 1]

    def attach_grads(self):
        if getattr(self, "is_replica", False):
            return
        """Make every LoRA parameter's .grad a view of the flat bucket (zeroing it when a grad
        was dropped by zero_grad(set_to_none=True))."""
        fresh = any(p.grad is None or p.grad.data_ptr() != self.grad_bucket.data_ptr() + 4 * off
                    for p, off, n in self.lora_views)
        if fresh:
            self.grad_bucket.zero_()
            for p, off, n in self.lora_views:
                p.grad = self.grad_bucket[off:off + n].view_as(p)

    def refresh_lora(self, full=True):
        """Rebuild the 16-bit LoRA operand images from the fp32 masters (after every optimiser step). full: the folded
        W_eff = W + (alpha/r) B A in both layouts plus the factor images; not full: the factor images only -- all a
        training step with lora_dropout > 0 reads (the folded images are then stale until the next full refresh)."""
        if full:
            N.check(self.L.cvflow_lora_refresh(self.handle, _stream()), "cvflow_lora_refresh")
            self._folded_stale = False
        else:
            N.check(self.L.cvflow_lora_refresh_factors(self.handle, _stream()), "cvflow_lora_refresh_factors")
            self._folded_stale = True
        self._merged_version = self._version()
        self._dirty = False

    def set_debug_dropout_mask(self, mask):
        """Explicit keep masks (uint8 [n_blocks][3][B*T][256], 1 = keep) instead of the hash RNG: parity tests only."""
        self._drop_mask = mask.contiguous() if mask is not None else None
        self._drop_active = -1.0      # force a re-send

    def dropout_seed(self):
        """Base seed of the mask hash: a function of torch.manual_seed (no RNG stream is consumed)."""
        return (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + 0x51ED270B) & (2 ** 63 - 1)

    def sync_dropout(self, training):
        """Tell the handle whether the LoRA dropout is active (the nn.Module's train()/eval() state decides)."""
        p = self.lora_dropout_p if (training and self.lora_r > 0) else 0.0
        if p != self._drop_active:
            seed = self.dropout_seed() if p > 0 else 0
            mk = self._drop_mask if p > 0 else None
            N.check(self.L.cvflow_set_lora_dropout(self.handle, float(p), C.c_uint64(seed),
                                                   C.c_void_p(mk.data_ptr()) if mk is not None else None,
                                                   int(mk.shape[2]) if mk is not None else 0), "cvflow_set_lora_dropout")
            self._drop_active = p
            self.ws_key = None       # the training workspace holds the u_d stashes only when dropout is on

    def dropout_seed_snapshot(self):
        """Current device-side seeds of the mask hash (this handle and its stream replicas)."""
        out = []
        for h in [self] + self.replicas:
            v = C.c_uint64(0)
            N.check(self.L.cvflow_lora_dropout_seed(h.handle, C.byref(v), None), "cvflow_lora_dropout_seed")
            out.append(int(v.value))
        return out

    def dropout_seed_restore(self, seeds):
        for h, s in zip([self] + self.replicas, seeds):
            v = C.c_uint64(int(s))
            N.check(self.L.cvflow_lora_dropout_seed(h.handle, None, C.byref(v)), "cvflow_lora_dropout_seed")

    def _version(self):
        try:
            return self.param_bucket._version
        except RuntimeError:      # inference tensor
            return -2

    def mark_dirty(self):
        self._dirty = True

    def sync_lora(self, need_folded=True):
        """Refresh the operand images when the LoRA parameters changed since the last refresh (torch in-place updates
        bump the bucket's version counter; raw-pointer updates call mark_dirty). need_folded: the coming forward reads
        the folded W_eff (eval(), or training with lora_dropout = 0)."""
        if self._dirty or self._version() != self._merged_version:
            self.refresh_lora(full=need_folded)
        elif need_folded and getattr(self, "_folded_stale", False):
            self.refresh_lora(full=True)

    def trains_unfolded(self):
        """True while training forwards run the un-folded LoRA branch (lora_dropout > 0 active on the handle)."""
        return self._drop_active != 0.0

    def check_trainable(self, est):
        """LoRA dropout follows the module's train()/eval() state, like nn.Dropout in the reference."""
        self.sync_dropout(bool(est.training) and not getattr(est, "cvflow_ignore_lora_dropout", False))

    # -- calls -----------------------------------------------------------------------------------
    def _workspace(self, B, T, training):
        key = (B, T, int(training))
        if self.ws_key != key:
            need = self.L.cvflow_workspace_bytes(self.handle, B, T, int(training))
            if need < 0:
                N.check(int(need), "cvflow_workspace_bytes")
            if self.ws is None or self.ws.numel() < need:
                self.ws = None
                self.ws = torch.empty(int(need), device=self.device, dtype=torch.uint8)
            N.check(self.L.cvflow_set_workspace(self.handle, C.c_void_p(self.ws.data_ptr()), self.ws.numel()),
                    "cvflow_set_workspace")
            self.ws_key = key

    def forward(self, x, mask, mu, t, spks, cond, keep=None, iso_len=0, training=False, B=None, out=None):
        T = x.shape[-1]
        B = B or max(x.shape[0], mu.shape[0])
        self._workspace(B, T, training)
        if out is None:
            out = torch.empty(B, 80, T, device=self.device, dtype=torch.float32)
        io = EstimatorIO()
        io.x, io.x_nb = x.data_ptr(), x.shape[0]
        io.mask, io.mask_nb = mask.data_ptr(), mask.shape[0]
        io.mu, io.mu_nb = mu.data_ptr(), mu.shape[0]
        io.t, io.t_nb = t.data_ptr(), t.shape[0]
        io.spks, io.spks_nb = (spks.data_ptr(), spks.shape[0]) if spks is not None else (None, 1)
        io.cond, io.cond_nb = (cond.data_ptr(), cond.shape[0]) if cond is not None else (None, 1)
        io.keep = keep.data_ptr() if keep is not None else None
        io.out = out.data_ptr()
        io.B, io.T, io.iso_len, io.training = B, T, int(iso_len), int(training)
        N.check(self.L.cvflow_estimator_forward(self.handle, C.byref(io), _stream()), "cvflow_estimator_forward")
        return out

    def time_embed(self, t):
        """time_mlp(SinusoidalPosEmb(t)) [B][1024] through cvflow_time_embed (modules.py:27-57)."""
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous().reshape(-1)
        B = t.shape[0]
        out = torch.empty(B, 1024, device=self.device, dtype=torch.float32)
        scratch = torch.empty(B * 1344, device=self.device, dtype=torch.float32)
        N.check(self.L.cvflow_time_embed(self.handle, t.data_ptr(), B, out.data_ptr(), scratch.data_ptr(), B, _stream()),
                "cvflow_time_embed")
        return out

    def backward(self, dpred16, grad_scale=1.0, grad_scale_dev=None, input_grads=None):
        """LoRA gradients into the flat bucket; `input_grads` = {'dx'|'dmu'|'dspks'|'dcond': fp32 tensor} additionally
        receives dL/d(estimator inputs) (for training the modules that produce mu / spks upstream)."""
        self.attach_grads()
        gs = C.c_void_p(grad_scale_dev.data_ptr()) if grad_scale_dev is not None else None
        if input_grads:
            ig = InputGrads()
            for k, t in input_grads.items():
                assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32, k
                setattr(ig, k, t.data_ptr())
            N.check(self.L.cvflow_estimator_backward_inputs(
                self.handle, C.c_void_p(dpred16.data_ptr()), float(grad_scale), gs, C.byref(ig), _stream()),
                "cvflow_estimator_backward_inputs")
            return
        N.check(self.L.cvflow_estimator_backward(
            self.handle, C.c_void_p(dpred16.data_ptr()), float(grad_scale), gs, _stream()),
            "cvflow_estimator_backward")

    def dpred_buffer(self, B, T):
        """Persistent dL/dpred buffer (its address is baked into a cached TMA descriptor)."""
        buf = getattr(self, "_dpred", None)
        if buf is None or buf.shape[0] != B or buf.shape[1] != T:
            buf = torch.empty(B, T, 128, device=self.device, dtype=self.dtype)
            self._dpred = buf
        return buf

    def launch_count(self):
        return int(self.L.cvflow_launch_count(self.handle))


def _prep(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _prep_g(t, device):
    """_prep that stays on the autograd graph when the tensor requires grad (inputs trained upstream)."""
    if t.requires_grad and torch.is_grad_enabled():
        return t.to(device=device, dtype=torch.float32).contiguous()
    return _prep(t, device)


def native_of(module, dtype=None):
    """Lazily create (or fetch) the NativeEstimator bound to `module`."""
    want = dtype or getattr(module, "cvflow_dtype", torch.float16)
    ne = module.__dict__.get("_cvflow")
    if ne is None or ne.dtype != want:
        ne = NativeEstimator(module, want)
        module.__dict__["_cvflow"] = ne
    return ne


class _EstimatorFn(torch.autograd.Function):
    """Generic autograd bridge for ConditionalDecoder.forward (any downstream loss). The CFM
    training step uses the fused path in flow_model.py instead."""

    @staticmethod
    def forward(ctx, ne, x, mask, mu, t, spks, cond, iso_len, *lora_params):
        ctx.ne = ne
        ctx.shapes = (x.shape, mu.shape, spks.shape if spks is not None else None,
                      cond.shape if cond is not None else None)
        out = ne.forward(x, mask, mu, t, spks, cond, iso_len=iso_len, training=True)
        return out

    @staticmethod
    def backward(ctx, gout):
        ne = ctx.ne
        B, _, T = gout.shape
        g = torch.zeros(B, T, 128, device=gout.device, dtype=ne.dtype)
        g[:, :, :80] = (gout.float() * ne.loss_scale).transpose(1, 2).to(ne.dtype)
        want = {}
        for key, pos, shape in (("dx", 1, ctx.shapes[0]), ("dmu", 3, ctx.shapes[1]), ("dspks", 5, ctx.shapes[2]),
                                ("dcond", 6, ctx.shapes[3])):
            if ctx.needs_input_grad[pos] and shape is not None:
                want[key] = (pos, torch.empty(shape, device=gout.device, dtype=torch.float32))
        ne.backward(g, grad_scale=1.0 / ne.loss_scale, input_grads={k: v for k, (_, v) in want.items()})
        grads = [None] * (8 + len(ne.lora_views))
        for _, (pos, v) in want.items():
            grads[pos] = v
        return tuple(grads)


def estimator_forward(module, x, mask, mu, t, spks=None, cond=None):
    ne = native_of(module)
    dev = ne.device
    x_, mu_, t_ = _prep_g(x, dev), _prep_g(mu, dev), _prep(t, dev)
    mask_ = _prep(mask, dev).reshape(mask.shape[0], -1)
    spks_ = _prep_g(spks, dev) if spks is not None else None
    cond_ = _prep_g(cond, dev) if cond is not None else None
    iso = int(module.prompt_isolation_len) if getattr(module, "prompt_isolation_enabled", False) else 0
    needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p, _, _ in ne.lora_views) or
                                              any(v is not None and v.requires_grad for v in (x_, mu_, spks_, cond_)))
    if needs_grad:
        ne.sync_dropout(bool(module.training) and not getattr(module, "cvflow_ignore_lora_dropout", False))
        ne.sync_lora(need_folded=not ne.trains_unfolded())
        out = _EstimatorFn.apply(ne, x_, mask_, mu_, t_, spks_, cond_, iso, *[p for p, _, _ in ne.lora_views])
    else:
        ne.sync_lora(need_folded=True)
        out = ne.forward(x_, mask_, mu_, t_, spks_, cond_, iso_len=iso, training=False)
    return out.to(x.dtype) if x.dtype != torch.float32 else out
