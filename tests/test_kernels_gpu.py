"""Unit parity of individual CUDA kernels that are reachable through the C ABI."""
import ctypes as C

import pytest
import torch

from cosyvoice_lora_finetune_framework_b200 import _estimator as E
from cosyvoice_lora_finetune_framework_b200 import _native as N

pytestmark = pytest.mark.gpu


def test_cfm_prep_and_loss():
    L = E._lib()
    torch.manual_seed(0)
    B, T = 3, 77
    x1, z = torch.randn(B, 80, T, device="cuda"), torch.randn(B, 80, T, device="cuda")
    t = torch.rand(B, device="cuda")
    y = torch.empty_like(x1)
    N.check(L.cvflow_cfm_prep(x1.data_ptr(), z.data_ptr(), t.data_ptr(), y.data_ptr(), B, T, 1e-6, E._stream()))
    ref = (1 - (1 - 1e-6) * t.view(B, 1, 1)) * z + t.view(B, 1, 1) * x1
    assert torch.allclose(y, ref, atol=1e-6)
    pred = torch.randn(B, 80, T, device="cuda")
    w = (torch.rand(B, T, device="cuda") > 0.3).float() * torch.tensor([1.0, 5.0, 1.0], device="cuda").view(B, 1)
    mask = (w > 0).float()
    scal = torch.zeros(4, device="cuda")
    partials = torch.zeros(B * ((T + 31) // 32), device="cuda")
    dp = torch.zeros(B, T, 128, device="cuda", dtype=torch.float16)
    N.check(L.cvflow_cfm_loss(pred.data_ptr(), x1.data_ptr(), z.data_ptr(), w.data_ptr(), mask.data_ptr(),
                              scal.data_ptr(), partials.data_ptr(), dp.data_ptr(), B, T, 1e-6, 1024.0, 0, None, E._stream()))
    p = pred.clone().requires_grad_(True)
    u = x1 - (1 - 1e-6) * z
    loss = (((p - u) * w[:, None]) ** 2).sum() / (w.sum() * 80)
    loss.backward()
    assert abs(scal[2].item() - loss.item()) <= 1e-5 * loss.item()
    g = dp[:, :, :80].float().transpose(1, 2) / 1024.0
    assert torch.allclose(g, p.grad * mask[:, None], atol=1e-3 * p.grad.abs().max().item())
    assert dp[:, :, 80:].abs().max().item() == 0


def test_euler_update_and_adamw():
    L = E._lib()
    torch.manual_seed(1)
    n = 80 * 123
    x = torch.randn(n, device="cuda")
    d = torch.randn(2 * n, device="cuda")
    dt = torch.tensor([0.1, 0.25], device="cuda")
    ref = x + 0.25 * (1.7 * d[:n] - 0.7 * d[n:])
    N.check(L.cvflow_euler_update(x.data_ptr(), d.data_ptr(), dt.data_ptr(), 1, 0.7, n, E._stream()))
    assert torch.allclose(x, ref, atol=1e-6)
    # fused clip + AdamW vs torch
    m = 10007
    p0 = torch.randn(m, device="cuda")
    g = torch.randn(m, device="cuda") * 3
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([pt], lr=1e-3, weight_decay=0.01)
    ours, mm, vv = p0.clone(), torch.zeros(m, device="cuda"), torch.zeros(m, device="cuda")
    partials, ss = torch.zeros(296, device="cuda"), torch.zeros(1, device="cuda")
    found = torch.zeros(1, device="cuda", dtype=torch.int32)
    for step in range(1, 4):
        pt.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([pt], 1.0)
        opt.step()
        N.check(L.cvflow_sumsq(g.data_ptr(), m, partials.data_ptr(), ss.data_ptr(), E._stream()))
        N.check(L.cvflow_adamw_step(ours.data_ptr(), g.data_ptr(), mm.data_ptr(), vv.data_ptr(), m, ss.data_ptr(), 1.0,
                                    1.0, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, found.data_ptr(), None, E._stream()))
    assert torch.allclose(ours, pt.data, atol=1e-6, rtol=1e-5)
    assert found.item() == 0
