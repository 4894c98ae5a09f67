"""Fused FeedForward kernel (both GEMMs + GELU + residual in one launch) and its backward against plain
fp32 PyTorch on the same 16-bit-rounded operands (reference modules.py:192-224,372-374)."""
import pytest
import torch
import torch.nn.functional as F

from cosyvoice_lora_finetune_framework_b200 import _estimator as E
from cosyvoice_lora_finetune_framework_b200 import _native as N

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [128, 200, 6400, 333 * 128 + 17])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_mlp_forward_backward(M, dt):
    L = E._lib()
    torch.manual_seed(M)
    x = (torch.randn(M, 256, device="cuda")).to(dt)
    w1 = (torch.randn(1024, 256, device="cuda") * 0.08).to(dt)
    w2 = (torch.randn(256, 1024, device="cuda") * 0.05).to(dt)
    b1, b2 = torch.randn(1024, device="cuda") * 0.3, torch.randn(256, device="cuda") * 0.3
    resid = torch.randn(M, 256, device="cuda")
    out = torch.full((M, 256), float("nan"), device="cuda")
    pre = torch.full((M, 1024), float("nan"), device="cuda", dtype=dt)
    N.check(L.cvflow_mlp_forward(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), resid.data_ptr(),
                                 out.data_ptr(), pre.data_ptr(), M, N.dtype_code(dt), 0, E._stream()))
    torch.cuda.synchronize()
    pre_ref = x.float() @ w1.float().t() + b1
    hid = F.gelu(pre_ref, approximate="tanh").to(dt).float()   # the kernel feeds the second GEMM 16-bit operands
    ref = resid + b2 + hid @ w2.float().t()
    tol = 2e-3 if dt == torch.float16 else 1.6e-2
    assert torch.isfinite(out).all() and torch.isfinite(pre.float()).all()
    assert (pre.float() - pre_ref).abs().max().item() <= tol * pre_ref.abs().max().item()
    assert (out - ref).abs().max().item() <= tol * ref.abs().max().item()

    # backward: dx = ((dy W2) o gelu'(pre)) W1 with the transposed weight images the estimator binds
    dy = torch.randn(M, 256, device="cuda").to(dt)
    w2_t, w1_t = w2.t().contiguous(), w1.t().contiguous()
    dx = torch.full((M, 256), float("nan"), device="cuda", dtype=dt)
    N.check(L.cvflow_mlp_backward(dy.data_ptr(), w2_t.data_ptr(), pre.data_ptr(), w1_t.data_ptr(), dx.data_ptr(), M,
                                  N.dtype_code(dt), 0, E._stream()))
    torch.cuda.synchronize()
    pr = pre.float().requires_grad_(True)
    F.gelu(pr, approximate="tanh").backward(dy.float() @ w2.float())
    dpre = pr.grad.to(dt).float()
    dx_ref = dpre @ w1.float()
    assert torch.isfinite(dx.float()).all()
    assert (dx.float() - dx_ref).abs().max().item() <= 2 * tol * dx_ref.abs().max().item()
