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
                                  ne.loss_scale, N.dtype_code(ne.dtype), st), "cvflow_cfm_loss")
        ctx.ne = ne
        ctx.dpred = dpred
        ctx.n_extra = len(lora_params)
        ctx.mark_non_differentiable(y)
        ne.last_pred = pred
        return scal[2].clone(), y

    @staticmethod
    def backward(ctx, gloss, gy):
        ne = ctx.ne
        g = gloss.detach().reshape(1).to(torch.float32).contiguous()
        ne.backward(ctx.dpred, grad_scale=1.0 / ne.loss_scale, grad_scale_dev=g)
        return (None,) * (12 + ctx.n_extra)


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
                      keep=torch.tensor([1.0, 0.0], device=dev), graph=None)
        st["x"].copy_(x.float())
        st["mu"].copy_(mu.float())
        st["mask"].copy_(mask.float().reshape(1, T))
        st["spks"].copy_(spks.float()) if spks is not None else st["spks"].zero_()
        st["cond"].copy_(cond.float()) if cond is not None else st["cond"].zero_()
        st["t"].copy_(t_arr)
        st["dt"].copy_(dt_arr)
        L = E._lib()

        def run():
            for k in range(n_steps):
                ne.forward(st["x"], st["mask"], st["mu"], st["t"][k:k + 1], st["spks"], st["cond"], keep=st["keep"],
                           iso_len=0, training=False, B=2, out=st["d"])
                N.check(L.cvflow_euler_update(st["x"].data_ptr(), st["d"].data_ptr(), st["dt"].data_ptr(), k,
                                              float(self.inference_cfg_rate), 80 * T, E._stream()), "cvflow_euler_update")

        if not self.use_cuda_graph:
            run()
        elif st["graph"] is None:
            x0 = st["x"].clone()
            run()                                  # warm-up: builds plans, workspace, function attributes
            torch.cuda.synchronize()
            st["x"].copy_(x0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run()
            st["graph"] = g
            self._graphs[key] = st
            st["x"].copy_(x0)
            g.replay()
        else:
            st["graph"].replay()
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
        spks_ = f(spks) if spks is not None else None
        cond_ = f(cond) if cond is not None else None
        keep_ = keep.to(dev).float().contiguous() if keep is not None else None
        ne.sync_lora()
        loss, y = _CFMLossFn.apply(ne, f(x1), f(mask).reshape(b, T), f(mu), spks_, cond_, f(w).reshape(b, T),
                                   f(t_step).reshape(b), f(z), keep_, iso, float(self.sigma_min),
                                   *[p for p, _, _ in ne.lora_views])
        return loss, y.to(x1.dtype)
