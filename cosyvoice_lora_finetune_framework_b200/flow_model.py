"""Conditional flow matching over the CUDA estimator (API mirror of the reference's
flow_model.py:50-204 `ConditionalCFM`).

  compute_loss  one fused training step: CFM interpolation kernel -> estimator forward (stashing)
                -> masked-loss + dL/dpred kernel; `loss.backward()` runs the estimator backward and
                leaves the LoRA gradients in a flat fp32 bucket (the DDP allreduce payload).
  forward       Euler ODE solve with classifier-free guidance; the whole N-step solve (2N+... kernel
                launches x N) is captured once per (T, N) into a CUDA graph and replayed.

Random draws use the same torch calls, shapes, dtypes, device and order as the reference
(rand([B,1,1]) -> randn_like(x1) -> rand(B); randn_like(mu)), so a seeded run consumes the RNG
stream identically.
"""
import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _estimator as E
from . import _native as N

try:
    from .config import ANTI_LEAKAGE_CONFIG
except ImportError:  # pragma: no cover
    ANTI_LEAKAGE_CONFIG = {'boundary_frames': 15, 'boundary_loss_weight': 3.0, 'boundary_loss_enabled': True}

_PI_HALF = 0.5 * 3.14159265359   # the reference's literal (flow_model.py:90,148)


class _CFMLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ne, x1, mask, mu, spks, cond, w, t, z, keep, iso_len, sigma_min, *lora_params):
        L = E._lib()
        B, _, T = x1.shape
        st = E._stream()
        y = torch.empty_like(x1)
        N.check(L.cvflow_cfm_prep(x1.data_ptr(), z.data_ptr(), t.data_ptr(), y.data_ptr(), B, T, sigma_min, st),
                "cvflow_cfm_prep")
        pred = ne.forward(y, mask, mu, t, spks, cond, keep=keep, iso_len=iso_len, training=True)
        scal = torch.zeros(4, device=x1.device, dtype=torch.float32)
        partials = torch.empty(B * ((T + 31) // 32), device=x1.device, dtype=torch.float32)
        dpred = ne.dpred_buffer(B, T)
        N.check(L.cvflow_cfm_loss(pred.data_ptr(), x1.data_ptr(), z.data_ptr(), w.data_ptr(), mask.data_ptr(),
                                  scal.data_ptr(), partials.data_ptr(), dpred.data_ptr(), B, T, sigma_min,
                                  ne.loss_scale, N.dtype_code(ne.dtype), None, st), "cvflow_cfm_loss")
        ctx.ne = ne
        ctx.dpred = dpred
        ctx.keep = keep          # the backward's unpack kernel reads the CFG keep factors again
        ctx.n_extra = len(lora_params)
        ctx.in_shapes = (mu.shape, spks.shape if spks is not None else None, cond.shape if cond is not None else None)
        ctx.mark_non_differentiable(y)
        ne.last_pred = pred
        return scal[2].clone(), y

    @staticmethod
    def backward(ctx, gloss, gy):
        ne = ctx.ne
        g = gloss.detach().reshape(1).to(torch.float32).contiguous()
        # dL/dmu, dL/dspks, dL/dcond only when the caller trains something upstream of the estimator
        want = {}
        for key, pos, shape in (("dmu", 3, ctx.in_shapes[0]), ("dspks", 4, ctx.in_shapes[1]), ("dcond", 5, ctx.in_shapes[2])):
            if ctx.needs_input_grad[pos] and shape is not None:
                want[key] = (pos, torch.empty(shape, device=g.device, dtype=torch.float32))
        ne.backward(ctx.dpred, grad_scale=1.0 / ne.loss_scale, grad_scale_dev=g,
                    input_grads={k: v for k, (_, v) in want.items()})
        grads = [None] * (12 + ctx.n_extra)
        for _, (pos, v) in want.items():
            grads[pos] = v
        return tuple(grads)


class _CFMLossShardedFn(torch.autograd.Function):
    """The same training step with the batch cut into contiguous shards that run CONCURRENTLY on
    several CUDA streams (one estimator handle per shard over shared weights). Every op of the path
    is per-utterance except the loss normaliser sum(w), which is computed once for the whole batch
    and handed to each shard, so the result equals the single-stream step. At these sizes
    (6,400-12,800 tokens x 256 channels per kernel) single kernels are latency-bound; concurrent
    shards fill the 148 SMs."""

    @staticmethod
    def forward(ctx, nes, streams, x1, mask, mu, spks, cond, w, t, z, keep, iso_len, sigma_min, *lora_params):
        from .parallel import shard_bounds
        L = E._lib()
        B, _, T = x1.shape
        S = len(nes)
        cur = torch.cuda.current_stream()
        wsum = w.sum().reshape(1)
        y = torch.empty_like(x1)
        scal = torch.zeros(S, 4, device=x1.device, dtype=torch.float32)
        shards = []
        for s in range(S):
            lo, hi = shard_bounds(B, s, S)
            st = cur if s == 0 else streams[s - 1]
            if s > 0:
                st.wait_stream(cur)
            with torch.cuda.stream(st):
                ne = nes[s]
                cs = C.c_void_p(st.cuda_stream)
                sl = slice(lo, hi)
                N.check(L.cvflow_cfm_prep(x1[sl].data_ptr(), z[sl].data_ptr(), t[sl].data_ptr(), y[sl].data_ptr(), hi - lo,
                                          T, sigma_min, cs), "cvflow_cfm_prep")
                pred = ne.forward(y[sl], mask[sl], mu[sl], t[sl], spks[sl] if spks is not None else None,
                                  cond[sl] if cond is not None else None, keep=keep[sl] if keep is not None else None,
                                  iso_len=iso_len, training=True)
                partials = torch.empty((hi - lo) * ((T + 31) // 32), device=x1.device, dtype=torch.float32)
                dpred = ne.dpred_buffer(hi - lo, T)
                N.check(L.cvflow_cfm_loss(pred.data_ptr(), x1[sl].data_ptr(), z[sl].data_ptr(), w[sl].data_ptr(),
                                          mask[sl].data_ptr(), scal[s].data_ptr(), partials.data_ptr(), dpred.data_ptr(),
                                          hi - lo, T, sigma_min, ne.loss_scale, N.dtype_code(ne.dtype), wsum.data_ptr(),
                                          cs), "cvflow_cfm_loss")
                shards.append((st, ne, dpred, pred, partials))
        for s in range(1, S):
            cur.wait_stream(streams[s - 1])
        denom = wsum * 80.0
        loss = torch.where(denom > 0, scal[:, 1].sum() / denom.clamp_min(1e-30), torch.zeros_like(denom)).reshape(())
        ctx.shards = shards
        ctx.streams = streams
        ctx.n_extra = len(lora_params)
        ctx.mark_non_differentiable(y)
        return loss, y

    @staticmethod
    def backward(ctx, gloss, gy):
        cur = torch.cuda.current_stream()
        g = gloss.detach().reshape(1).to(torch.float32).contiguous()
        primary = ctx.shards[0][1]
        primary.attach_grads()
        for s, (st, ne, dpred, _, _) in enumerate(ctx.shards):
            if s > 0:
                st.wait_stream(cur)
            with torch.cuda.stream(st):
                ne.backward(dpred, grad_scale=1.0 / ne.loss_scale, grad_scale_dev=g)
        for s in range(1, len(ctx.shards)):
            cur.wait_stream(ctx.shards[s][0])
        for s in range(1, len(ctx.shards)):        # deterministic, fixed-order sum of the shard buckets
            rb = ctx.shards[s][1].grad_bucket
            primary.grad_bucket.add_(rb)
            rb.zero_()
        return (None,) * (13 + ctx.n_extra)


class ConditionalCFM(nn.Module):
    """Conditional Flow Matching module (same constructor / methods as the reference)."""

    def __init__(self, in_channels: int, n_spks: int = 1, spk_emb_dim: int = 64, sigma_min: float = 1e-6,
                 t_scheduler: str = 'cosine', training_cfg_rate: float = 0.2, inference_cfg_rate: float = 0.7,
                 estimator: Optional[nn.Module] = None):
        super().__init__()
        self.in_channels = in_channels
        self.n_spks = n_spks
        self.spk_emb_dim = spk_emb_dim
        self.sigma_min = sigma_min
        self.t_scheduler = t_scheduler
        self.training_cfg_rate = training_cfg_rate
        self.inference_cfg_rate = inference_cfg_rate
        self.estimator = estimator
        self.use_cuda_graph = True
        self._graphs = {}
        self.num_streams = 1      # >1: training batch shards run concurrently on that many CUDA streams
        self._streams = []

    # ------------------------------------------------------------------------------------------
    # inference
    # ------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, prompt_len=0, cache=None):
        z = torch.randn_like(mu).to(mu.device).to(mu.dtype) * temperature
        return self._forward_with_noise(z, mu, mask, n_timesteps, spks, cond, prompt_len, cache)

    def _forward_with_noise(self, z, mu, mask, n_timesteps, spks=None, cond=None, prompt_len=0, cache=None):
        """forward() with z (already scaled by the temperature) supplied: used by the parity tests."""
        if cache is not None and cache.shape[2] != 0:
            n = cache.shape[2]
            z[:, :, :n] = cache[:, :, :, 0]
            mu[:, :, :n] = cache[:, :, :, 1]      # in place, like the reference (flow_model.py:82)
        if prompt_len > 0:
            z_cache = torch.concat([z[:, :, :prompt_len], z[:, :, -34:]], dim=2)
            mu_cache = torch.concat([mu[:, :, :prompt_len], mu[:, :, -34:]], dim=2)
        else:
            z_cache, mu_cache = z[:, :, -34:], mu[:, :, -34:]
        new_cache = torch.stack([z_cache, mu_cache], dim=-1)
        t_span = torch.linspace(0, 1, n_timesteps + 1, device=mu.device, dtype=mu.dtype)
        if self.t_scheduler == 'cosine':
            t_span = 1 - torch.cos(t_span * _PI_HALF)
        return self.solve_euler(z, t_span, mu, mask, spks, cond), new_cache

    def _time_grid(self, t_span):
        """(t_k, dt_k) exactly as the reference accumulates them (flow_model.py:96,120-123)."""
        t_span = t_span.float()
        t = t_span[0]
        dt = t_span[1] - t_span[0]
        ts, dts = [], []
        n = t_span.shape[0]
        for step in range(1, n):
            ts.append(t)
            dts.append(dt)
            t = t + dt
            if step < n - 1:
                dt = t_span[step + 1] - t
        return torch.stack(ts).contiguous(), torch.stack(dts).contiguous()

    def solve_euler(self, x, t_span, mu, mask, spks, cond):
        """Fixed-step Euler with CFG (flow_model.py:94-125). Batch 1 only, like the reference."""
        assert self.estimator is not None
        if x.shape[0] != 1:
            raise ValueError("solve_euler packs cond/uncond into batch 2 and therefore needs batch 1 "
                             "(reference flow_model.py:100-105)")
        ne = E.native_of(self.estimator)
        dev = ne.device
        T = x.shape[2]
        n_steps = t_span.shape[0] - 1
        t_arr, dt_arr = self._time_grid(t_span.to(dev))
        key = (T, n_steps, ne.dtype)
        st = self._graphs.get(key) if self.use_cuda_graph else None
        if st is None:
            st = dict(x=torch.empty(1, 80, T, device=dev), mu=torch.empty(1, 80, T, device=dev),
                      mask=torch.empty(1, T, device=dev), spks=torch.zeros(1, 80, device=dev),
                      cond=torch.zeros(1, 80, T, device=dev), t=torch.empty(n_steps, device=dev),
                      dt=torch.empty(n_steps, device=dev), d=torch.empty(2, 80, T, device=dev),
                      keep=torch.tensor([1.0, 0.0], device=dev))
        st["x"].copy_(x.float())
        st["mu"].copy_(mu.float())
        st["mask"].copy_(mask.float().reshape(1, T))
        st["spks"].copy_(spks.float()) if spks is not None else st["spks"].zero_()
        st["cond"].copy_(cond.float()) if cond is not None else st["cond"].zero_()
        st["t"].copy_(t_arr)
        st["dt"].copy_(dt_arr)
        L = E._lib()

        def run():      # eager form: 2 launches of the C ABI per step
            for k in range(n_steps):
                ne.forward(st["x"], st["mask"], st["mu"], st["t"][k:k + 1], st["spks"], st["cond"], keep=st["keep"],
                           iso_len=0, training=False, B=2, out=st["d"])
                N.check(L.cvflow_euler_update(st["x"].data_ptr(), st["d"].data_ptr(), st["dt"].data_ptr(), k,
                                              float(self.inference_cfg_rate), 80 * T, E._stream()), "cvflow_euler_update")

        if not self.use_cuda_graph:
            run()
        else:
            # The whole N-step solve is ONE CUDA graph built and owned by the library (cvflow_solve_capture /
            # cvflow_solve_replay): any binder of the C ABI gets the same graph-replayed solve, no Python in the loop.
            # The handle keeps one captured solve per (T, n_steps); they share the workspace arena (every solve writes
            # what it reads), so they stay valid until the arena is re-allocated.
            ne.sync_lora()
            ne._workspace(2, T, False)
            if getattr(ne, "_solve_ws", None) != ne.ws.data_ptr():
                N.check(L.cvflow_solve_release(ne.handle), "cvflow_solve_release")
                ne._solve_ws, ne._solve_keys = ne.ws.data_ptr(), {}
            if ne._solve_keys.get((T, n_steps)) != id(st):
                x0 = st["x"].clone()
                cap = torch.cuda.Stream(device=dev)            # capture needs a non-default stream
                cap.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cap):
                    N.check(L.cvflow_solve_capture(ne.handle, T, n_steps, float(self.inference_cfg_rate), st["x"].data_ptr(),
                                                   st["mask"].data_ptr(), st["mu"].data_ptr(), st["spks"].data_ptr(),
                                                   st["cond"].data_ptr(), st["t"].data_ptr(), st["dt"].data_ptr(),
                                                   st["d"].data_ptr(), C.c_void_p(cap.cuda_stream)), "cvflow_solve_capture")
                torch.cuda.current_stream().wait_stream(cap)
                st["x"].copy_(x0)
                ne._solve_keys[(T, n_steps)] = id(st)
                self._graphs[key] = st
            N.check(L.cvflow_solve_replay(ne.handle, T, n_steps, E._stream()), "cvflow_solve_replay")
        return st["x"].clone().float()

    # ------------------------------------------------------------------------------------------
    # training
    # ------------------------------------------------------------------------------------------
    def compute_loss(self, x1, mask, mu, spks=None, cond=None, prompt_lens=None):
        """CFM loss with prompt masking / boundary weights and prompt isolation
        (flow_model.py:127-204). Returns (loss, y)."""
        assert self.estimator is not None
        b = mu.shape[0]
        t_step = torch.rand([b, 1, 1], device=mu.device, dtype=mu.dtype)
        if self.t_scheduler == 'cosine':
            t_step = 1 - torch.cos(t_step * _PI_HALF)
        z = torch.randn_like(x1)
        keep = None
        if self.training_cfg_rate > 0:
            keep = torch.rand(b, device=x1.device) > self.training_cfg_rate
        return self._loss_with_noise(x1, mask, mu, spks, cond, prompt_lens, t_step, z, keep)

    def _loss_with_noise(self, x1, mask, mu, spks, cond, prompt_lens, t_step, z, keep):
        """compute_loss with the random draws supplied (t already warped): used by the parity tests."""
        est = self.estimator
        if torch.is_tensor(x1) and x1.requires_grad:
            raise NotImplementedError("compute_loss: x1 (the target mel) requires grad; the CUDA flow path produces "
                                      "dL/dmu, dL/dspks and dL/dcond, not dL/dx1")
        upstream = torch.is_grad_enabled() and any(torch.is_tensor(v) and v.requires_grad for v in (mu, spks, cond))
        ne = E.native_of(est)
        ne.check_trainable(est)
        dev = ne.device
        iso = 0
        if prompt_lens is not None and len(prompt_lens) > 0:
            iso = int(max(prompt_lens))
            est.prompt_isolation_enabled = True
        iso = iso if getattr(est, 'prompt_isolation_enabled', False) else 0
        est.prompt_isolation_len = 0
        w = mask.clone()
        if prompt_lens is not None:
            frames = ANTI_LEAKAGE_CONFIG.get('boundary_frames', 15)
            weight = ANTI_LEAKAGE_CONFIG.get('boundary_loss_weight', 3.0)
            for i, p in enumerate(prompt_lens):
                if p > 0:
                    w[i, :, :p] = 0
                    if ANTI_LEAKAGE_CONFIG.get('boundary_loss_enabled', True):
                        w[i, :, p:min(p + frames, w.shape[2])] = weight
        b, _, T = x1.shape
        f = lambda v: E._prep(v, dev)
        fg = lambda v: E._prep_g(v, dev)
        spks_ = fg(spks) if spks is not None else None
        cond_ = fg(cond) if cond is not None else None
        keep_ = keep.to(dev).float().contiguous() if keep is not None else None
        ne.sync_lora(need_folded=not ne.trains_unfolded())
        S = min(int(self.num_streams), b)
        if S > 1:
            if upstream:
                raise NotImplementedError("num_streams > 1 does not produce dL/d(mu, spks, cond); use num_streams = 1")
            if ne._drop_active != 0.0:
                raise NotImplementedError("num_streams > 1 does not support lora_dropout > 0; use num_streams = 1")
            while len(self._streams) < S - 1:
                self._streams.append(torch.cuda.Stream(device=dev))
            loss, y = _CFMLossShardedFn.apply(ne.shard_handles(S), self._streams[: S - 1], f(x1), f(mask).reshape(b, T),
                                              f(mu), spks_, cond_, f(w).reshape(b, T), f(t_step).reshape(b), f(z),
                                              keep_, iso, float(self.sigma_min), *[p for p, _, _ in ne.lora_views])
        else:
            loss, y = _CFMLossFn.apply(ne, f(x1), f(mask).reshape(b, T), fg(mu), spks_, cond_, f(w).reshape(b, T),
                                       f(t_step).reshape(b), f(z), keep_, iso, float(self.sigma_min),
                                       *[p for p, _, _ in ne.lora_views])
        return loss, y.to(x1.dtype)


# ==============================================================================================
# Callers of the CUDA path: the flow model wrapper that prepares (x1, mask, mu, spks, cond) and
# the model builder (reference flow_model.py:207-767). Host-side PyTorch; the conditional flow
# matching itself runs through ConditionalCFM above.
# ==============================================================================================
import random  # noqa: E402
from typing import Any, Dict  # noqa: E402

import torch.nn.functional as F  # noqa: E402

from .encoder import ConformerEncoder, InterpolateRegulator  # noqa: E402
from .modules import ConditionalDecoder  # noqa: E402
from .utils import make_pad_mask  # noqa: E402

try:
    from .config import MEL_MEAN, MEL_STD, NO_PROMPT_TRAINING_CONFIG
except ImportError:  # pragma: no cover
    MEL_MEAN, MEL_STD = -6.0, 2.0
    NO_PROMPT_TRAINING_CONFIG = {'enabled': False, 'mode': 'full', 'no_prompt_ratio': 0.8, 'use_mean_embedding': False}


def _ode_steps_for(total_mel_len: int) -> int:
    """10 / 15 / 20 Euler steps for <=300 / <=500 / >500 frames (flow_model.py:530-536)."""
    if total_mel_len > 500:
        return 20
    if total_mel_len > 300:
        return 15
    return 10


class MaskedDiffWithXvec(nn.Module):
    """Flow model wrapper: token embedding -> Conformer encoder -> length regulator -> CFM decoder."""

    def __init__(self, input_size: int = 512, output_size: int = 80, spk_embed_dim: int = 192, vocab_size: int = 4096,
                 input_frame_rate: int = 50, encoder: Optional[nn.Module] = None,
                 length_regulator: Optional[nn.Module] = None, decoder: Optional[nn.Module] = None):
        super().__init__()
        self.input_size = input_size
        self.output_size = output_size
        self.vocab_size = vocab_size
        self.input_frame_rate = input_frame_rate
        self.input_embedding = nn.Embedding(vocab_size, input_size)
        self.spk_embed_affine_layer = nn.Linear(spk_embed_dim, output_size)
        self.encoder = encoder
        assert self.encoder is not None
        self.encoder_proj = nn.Linear(self.encoder.output_size(), output_size)
        self.decoder = decoder
        self.length_regulator = length_regulator
        self.mel_mean = MEL_MEAN
        self.mel_std = MEL_STD

    def normalize_mel(self, mel: torch.Tensor) -> torch.Tensor:
        return (mel - self.mel_mean) / self.mel_std

    def denormalize_mel(self, mel: torch.Tensor) -> torch.Tensor:
        return mel * self.mel_std + self.mel_mean

    # -- shared pieces -----------------------------------------------------------------------------
    def _speaker(self, embedding):
        return self.spk_embed_affine_layer(F.normalize(embedding, dim=1))

    # Mixed precision of the host-side encoder on a CUDA device: None = fp32 (exact, the parity tests), or torch.bfloat16 /
    # torch.float16 = torch.autocast around the Conformer encoder, which is what the reference's trainer does to the whole
    # model (Lightning precision '16-mixed', config.py:76; train_joint.py sets this to the estimator's operand dtype).
    encoder_autocast = None

    # The host-side encoder is ~1,500 small ATen launches per step (forward + backward), i.e. launch-bound. With
    # encoder_cuda_graphs = True its forward and backward are captured once per input shape
    # (torch.cuda.make_graphed_callables) and replayed; meant for pipelines that pad batches to a few fixed shapes.
    encoder_cuda_graphs = False
    _enc_graph_cache_limit = 8

    def _encode(self, token, token_len, like):
        keep = (~make_pad_mask(token_len)).unsqueeze(-1).to(like)
        emb = self.input_embedding(torch.clamp(token, min=0)) * keep
        if self.encoder_cuda_graphs and token.is_cuda and torch.is_grad_enabled() and self.training:
            return self._encode_graphed(emb, token_len)
        if self.encoder_autocast is not None and token.is_cuda:
            with torch.autocast('cuda', dtype=self.encoder_autocast):
                h, _ = self.encoder(emb, token_len)
                return self.encoder_proj(h).float()
        h, _ = self.encoder(emb, token_len)
        return self.encoder_proj(h)

    def _encode_graphed(self, emb, token_len):
        cache = self.__dict__.setdefault("_enc_graphs", {})
        key = (tuple(emb.shape), self.encoder_autocast, self.encoder.training)
        fn = cache.get(key)
        ac = self.encoder_autocast

        class _Enc(nn.Module):
            def __init__(s, enc, proj):
                super().__init__()
                s.enc, s.proj = enc, proj

            def forward(s, x, n):
                h, _ = s.enc(x, n)
                return s.proj(h).float()

        if fn is None:
            if len(cache) >= self._enc_graph_cache_limit:
                cache.pop(next(iter(cache)))
            mod = _Enc(self.encoder, self.encoder_proj)
            sample = (emb.detach().clone(), token_len.clone())
            if ac is not None:
                with torch.autocast('cuda', dtype=ac, cache_enabled=False):
                    fn = torch.cuda.make_graphed_callables(mod, sample)
            else:
                fn = torch.cuda.make_graphed_callables(mod, sample)
            cache[key] = fn
        if ac is not None:
            with torch.autocast('cuda', dtype=ac, cache_enabled=False):
                return fn(emb, token_len)
        return fn(emb, token_len)

    def _loss(self, feat, feat_len, h, embedding, conds, prompt_lens):
        mask = (~make_pad_mask(feat_len)).to(h)
        loss, _ = self.decoder.compute_loss(feat.transpose(1, 2).contiguous(), mask.unsqueeze(1),
                                            h.transpose(1, 2).contiguous(), embedding, cond=conds.transpose(1, 2),
                                            prompt_lens=prompt_lens)
        return {'loss': loss}

    # -- training ----------------------------------------------------------------------------------
    def _prompt_plan(self, lens, cross_lens, have_cross):
        """The anti-leakage decisions of the reference's per-utterance loop (flow_model.py:319-387) from HOST lengths:
        one (prompt frames, silence-gap frames, prompt from the cross sample, recorded prompt_len, blind the text side)
        tuple per utterance. Python's `random` is consumed in the reference's order."""
        from .config import ANTI_LEAKAGE_CONFIG as AL
        silence_on = AL.get('silence_padding_enabled', False)
        dynamic_on = AL.get('dynamic_prompt_enabled', True)
        dropout_on = AL.get('prompt_dropout_enabled', True)
        blind_on = AL.get('text_blinding_enabled', True)
        cross_on = AL.get('cross_sample_enabled', True)
        lo_ratio, hi_ratio = AL.get('prompt_min_ratio', 0.10), AL.get('prompt_max_ratio', 0.30)
        p_drop, p_blind = AL.get('prompt_dropout_prob', 0.10), AL.get('text_blinding_prob', 0.7)
        sil_lo, sil_hi = AL.get('silence_min_tokens', 5), AL.get('silence_max_tokens', 10)
        plan = []
        for i, n in enumerate(lens):
            if dropout_on and random.random() < p_drop:                 # strategy 3: prompt dropout
                plan.append((0, 0, False, 0, False))
                continue
            if dynamic_on:                                              # strategy 2: dynamic prompt length
                lo = max(1, int(lo_ratio * n))
                p = random.randint(lo, max(lo + 1, int(hi_ratio * n)))
            else:
                p = max(1, int(0.3 * n))
            use_cross = bool(cross_on and have_cross and cross_lens is not None and cross_lens[i] > 0)
            if use_cross:                                               # strategy 5: cross-sample prompt
                p = min(p, int(cross_lens[i]))
            gap = 0
            if silence_on:                                              # strategy 1: silence gap
                g = int(random.randint(sil_lo, sil_hi) * 22050 / 256 / self.input_frame_rate)
                g = max(3, min(g, 20))
                if p + g < n:
                    gap = g
            blind = bool(blind_on and random.random() < p_blind)        # strategy 6: text-side blinding
            plan.append((p, gap, use_cross, p + gap, blind))
        return plan

    def _no_prompt_plan(self, lens):
        mode = NO_PROMPT_TRAINING_CONFIG.get('mode', 'full')
        ratio = NO_PROMPT_TRAINING_CONFIG.get('no_prompt_ratio', 0.8)
        plan = []
        for n in lens:
            if mode == 'full' or random.random() < ratio:
                plan.append((0, 0, False, 0, False))
            else:
                p = random.randint(1, max(2, int(0.1 * n)))
                plan.append((p, 0, False, p, False))
        return plan

    @staticmethod
    def _host_ints(t):
        """Lengths as Python ints. The collate function leaves them on the host (no device synchronisation); a device
        tensor costs one."""
        return [int(v) for v in (t.tolist() if torch.is_tensor(t) else t)]

    def forward(self, batch: dict, device: torch.device) -> Dict[str, Any]:
        """Training forward with the reference's anti-leakage prompt strategies
        (flow_model.py:248-400). Python's `random` is consumed in the same order as the reference. On a CUDA device
        everything in front of compute_loss except the Conformer encoder runs in csrc/regulator.cu (speaker affine,
        length regulator with its autograd, mel normalisation + conditioning + mask in one launch) and the loop makes no
        device synchronisation; on the CPU (host-logic tests) the same plan is applied with torch ops."""
        from .config import ANTI_LEAKAGE_CONFIG as AL
        device = torch.device(device)
        dtype = self.input_embedding.weight.dtype
        lens = self._host_ints(batch['speech_feat_len'])
        no_prompt = NO_PROMPT_TRAINING_CONFIG.get('enabled', False)
        cross_raw, cross_lens = None, None
        if not no_prompt and 'cross_sample_mel' in batch:
            cross_raw = batch['cross_sample_mel'].to(device).to(dtype)
            if batch.get('cross_sample_mel_len', None) is not None:
                cross_lens = self._host_ints(batch['cross_sample_mel_len'])
        token = batch['speech_token'].to(device)
        token_len = batch['speech_token_len'].to(device)
        feat_raw = batch['speech_feat'].to(device).to(dtype)
        feat_len = batch['speech_feat_len'].to(device)
        embedding = batch['embedding'].to(device).to(dtype)
        silence_value = (AL.get('silence_mel_value', -11.5) - self.mel_mean) / self.mel_std

        if device.type == 'cuda':
            from . import _path_inputs as PI
            if feat_raw.shape[1] != max(lens):
                raise ValueError("speech_feat must be padded to max(speech_feat_len) (the length regulator produces "
                                 "max(speech_feat_len) frames, modules.py:822)")
            spks = PI.spk_affine(self.spk_embed_affine_layer, embedding)
            h = self._encode(token, token_len, torch.zeros((), dtype=dtype, device=device))
            plan = self._no_prompt_plan(lens) if no_prompt else self._prompt_plan(lens, cross_lens, cross_raw is not None)
            desc = torch.tensor([[n, p, gap, 1 if xs else 0] for n, (p, gap, xs, _, _) in zip(lens, plan)], dtype=torch.int32)
            blind = torch.tensor([p if b else 0 for (p, _, _, _, b) in plan], dtype=torch.int32)
            both = torch.cat([desc.reshape(-1), blind]).pin_memory().to(device, non_blocking=True)
            B = len(lens)
            mu = PI.regulate(self.length_regulator, h, feat_raw.shape[1], lens=both[:4 * B:4], blind=both[4 * B:],
                             channel_major=True)
            x1, cond, mask = PI.pack_inputs(feat_raw, cross_raw, both[:4 * B], self.mel_mean, self.mel_std, silence_value)
            self.path_inputs_backend = "libcvflow (regulator.cu)"
            loss, _ = self.decoder.compute_loss(x1, mask, mu, spks, cond=cond, prompt_lens=[r for (_, _, _, r, _) in plan])
            return {'loss': loss}

        feat = self.normalize_mel(feat_raw)
        cross_mel = self.normalize_mel(cross_raw) if cross_raw is not None else None
        embedding = self._speaker(embedding)
        h = self._encode(token, token_len, torch.zeros((), dtype=dtype, device=device))
        h, _ = self.length_regulator(h, feat_len)
        plan = self._no_prompt_plan(lens) if no_prompt else self._prompt_plan(lens, cross_lens, cross_mel is not None)
        conds = torch.zeros(feat.shape, device=device, dtype=dtype)
        for i, (p, gap, use_cross, _, blind) in enumerate(plan):
            if p > 0:
                conds[i, :p] = (cross_mel if use_cross else feat)[i, :p]
            if gap > 0:
                conds[i, p:p + gap] = silence_value
            if blind:
                h[i, :p, :] = 0.0
        return self._loss(feat, feat_len, h, embedding, conds, [r for (_, _, _, r, _) in plan])

    # -- inference ---------------------------------------------------------------------------------
    @torch.inference_mode()
    def inference(self, token, token_len, prompt_token, prompt_token_len, prompt_feat, prompt_feat_len, embedding,
                  flow_cache=None):
        """Prompted inference (flow_model.py:475-551); no mel (de)normalisation, like upstream."""
        assert token.shape[0] == 1
        embedding = self._speaker(embedding)
        n_prompt, n_target = prompt_token.shape[1], token.shape[1]
        all_tokens = torch.concat([prompt_token, token], dim=1)
        h = self._encode(all_tokens, prompt_token_len + token_len, embedding)
        mel_len1 = prompt_feat.shape[1]
        mel_len2 = int(n_target / self.input_frame_rate * 22050 / 256)
        if hasattr(self.length_regulator, 'inference'):
            h, _ = self.length_regulator.inference(h[:, :n_prompt], h[:, n_prompt:], mel_len1, mel_len2,
                                                   self.input_frame_rate)
        else:
            h, _ = self.length_regulator(h, torch.tensor([mel_len1 + mel_len2], device=token.device))
        total = mel_len1 + mel_len2
        conds = torch.zeros([1, total, self.output_size], device=token.device).to(h.dtype)
        conds[:, :mel_len1] = prompt_feat
        mask = (~make_pad_mask(torch.tensor([total], device=token.device))).to(h)
        feat, new_cache = self.decoder(mu=h.transpose(1, 2).contiguous(), mask=mask.unsqueeze(1), spks=embedding,
                                       cond=conds.transpose(1, 2), n_timesteps=_ode_steps_for(total),
                                       prompt_len=mel_len1, cache=flow_cache)
        return feat[:, :, mel_len1:].float(), new_cache

    @torch.inference_mode()
    def inference_like_training(self, token, token_len, feat_len, embedding, prompt_feat=None, prompt_len=0,
                                n_timesteps=10):
        """Inference that mirrors the training layout: full token sequence, optional short prompt
        (flow_model.py:553-638). Returns the whole mel, prompt region included."""
        assert token.shape[0] == 1
        n = int(feat_len.item()) if isinstance(feat_len, torch.Tensor) else int(feat_len)
        embedding = self._speaker(embedding)
        h = self._encode(token, token_len, embedding)
        h, _ = self.length_regulator(h, torch.tensor([n], device=token.device))
        conds = torch.zeros([1, n, self.output_size], device=token.device, dtype=h.dtype)
        if prompt_feat is not None and prompt_len > 0:
            k = min(prompt_len, prompt_feat.shape[1], n)
            conds[:, :k] = prompt_feat[:, :k]
        if n_timesteps is None or n_timesteps == 10:
            n_timesteps = _ode_steps_for(n)
        mask = torch.ones([1, 1, n], device=token.device, dtype=h.dtype)
        feat, _ = self.decoder(mu=h.transpose(1, 2).contiguous(), mask=mask, spks=embedding, cond=conds.transpose(1, 2),
                               n_timesteps=n_timesteps, prompt_len=prompt_len if prompt_feat is not None else 0,
                               cache=None)
        return feat.float()


def build_flow_model(pretrained_path: Optional[str] = None, device: str = 'cuda', input_size: int = 512,
                     output_size: int = 80, spk_embed_dim: int = 192, vocab_size: int = 4096,
                     encoder_attention_heads: int = 8, encoder_linear_units: int = 2048, encoder_num_blocks: int = 6,
                     decoder_channels: tuple = (256, 256), decoder_attention_head_dim: int = 64,
                     decoder_n_blocks: int = 4, decoder_num_mid_blocks: int = 12,
                     decoder_num_heads: int = 8) -> MaskedDiffWithXvec:
    """CosyVoice-300M flow model (reference flow_model.py:641-767): same hyper-parameters, same
    construction order (hence the same seeded init), strict `flow.pt` load with the reference's
    shape-matched partial fallback."""
    import os
    encoder = ConformerEncoder(input_size=input_size, output_size=input_size, attention_heads=encoder_attention_heads,
                               linear_units=encoder_linear_units, num_blocks=encoder_num_blocks, dropout_rate=0.1,
                               positional_dropout_rate=0.1, attention_dropout_rate=0.1, normalize_before=True,
                               cnn_module_kernel=15, use_cnn_module=False, macaron_style=False, causal=False)
    length_regulator = InterpolateRegulator(channels=output_size, sampling_ratios=(1, 1, 1, 1),
                                            out_channels=output_size, groups=1)
    estimator = ConditionalDecoder(in_channels=320, out_channels=80, channels=decoder_channels, dropout=0.0,
                                   attention_head_dim=decoder_attention_head_dim, n_blocks=decoder_n_blocks,
                                   num_mid_blocks=decoder_num_mid_blocks, num_heads=decoder_num_heads, act_fn='gelu')
    decoder = ConditionalCFM(in_channels=output_size, n_spks=1, spk_emb_dim=output_size, sigma_min=1e-6,
                             t_scheduler='cosine', training_cfg_rate=0.2, inference_cfg_rate=0.7, estimator=estimator)
    model = MaskedDiffWithXvec(input_size=input_size, output_size=output_size, spk_embed_dim=spk_embed_dim,
                               vocab_size=vocab_size, input_frame_rate=50, encoder=encoder,
                               length_regulator=length_regulator, decoder=decoder)
    if pretrained_path is not None:
        weight_file = os.path.join(pretrained_path, 'flow.pt') if os.path.isdir(pretrained_path) else pretrained_path
        if os.path.exists(weight_file):
            print(f"Loading pretrained weights from: {weight_file}")
            state = torch.load(weight_file, map_location='cpu')
            try:
                model.load_state_dict(state, strict=True)
                print("Weights loaded successfully (strict=True)")
            except Exception as e:
                print(f"Strict loading failed: {e}")
                own = model.state_dict()
                matched = {k: v for k, v in state.items() if k in own and own[k].shape == v.shape}
                own.update(matched)
                model.load_state_dict(own, strict=False)
                print(f"Partial loading: {len(matched)}/{len(state)} weights loaded")
        else:
            print(f"Warning: Weight file not found: {weight_file}")
            print("Using random initialization")
    return model.to(device)
