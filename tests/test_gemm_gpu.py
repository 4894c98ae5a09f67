"""tcgen05 implicit-GEMM engine vs plain PyTorch fp32 on the same 16-bit-rounded operands."""
import ctypes as C

import pytest
import torch

from cosyvoice_lora_finetune_framework_b200 import _native as N

pytestmark = pytest.mark.gpu


def _desc(A, W, out, *, segs, R, nbatch=1, a_rows=None, dtype=torch.float16, **kw):
    d = N.GemmDesc()
    srcs = A if isinstance(A, (list, tuple)) else [A]
    for i, a in enumerate(srcs):
        d.A[i] = a.data_ptr()
        d.a_rows[i] = a.shape[-2] if a_rows is None else a_rows[i]
        d.a_cols[i] = a.shape[-1]
        d.a_ld[i] = a.stride(-2)
        d.a_bstride[i] = a.stride(0) if a.dim() == 3 else a.shape[-2] * a.stride(-2)
    d.nbatch = nbatch
    d.dtype = N.dtype_code(dtype)
    d.W = W.data_ptr()
    d.N = W.shape[0]
    d.Ktot = W.shape[1]
    for i, (m, sh, c0, nkb) in enumerate(segs):
        d.seg[i] = N.GemmSeg(m, sh, c0, nkb)
    d.nseg = len(segs)
    d.R = R
    d.rmul = kw.get("rmul", 1)
    d.roff = kw.get("roff", 0)
    d.out_rows = kw.get("out_rows", R)
    d.out = out.data_ptr()
    d.out_f32 = int(out.dtype == torch.float32)
    d.transposed_out = kw.get("transposed_out", 0)
    d.ldc = kw.get("ldc", out.shape[-1])
    d.col_off = kw.get("col_off", 0)
    d.n_valid = kw.get("n_valid", W.shape[0])
    d.alpha = kw.get("alpha", 1.0)
    d.act = kw.get("act", 0)
    for name in ("bias", "aux_out", "mul_src", "rowmask", "resid", "gn_part", "ln_gamma", "ln_beta"):
        t = kw.get(name)
        setattr(d, name, t.data_ptr() if t is not None else None)
    d.ld_aux = kw.get("ld_aux", 0)
    d.ldr = kw.get("ldr", 0)
    return d


def _run(d):
    N.check(N.lib().cvflow_gemm(C.byref(d), N.current_stream()), "cvflow_gemm")
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,Nn,K", [(128, 128, 64), (128, 128, 256), (400, 1536, 256), (1000, 256, 1024),
                                    (6400, 256, 512), (77, 256, 256), (40000, 1536, 256)])
def test_linear(M, Nn, K, dtype):
    torch.manual_seed(0)
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    W = (torch.randn(Nn, K, device="cuda") * 0.1).to(dtype)
    out = torch.full((M, Nn), float("nan"), device="cuda", dtype=dtype)
    d = _desc(A, W, out, segs=[(0, 0, 0, K // 64)], R=M, dtype=dtype)
    _run(d)
    ref = A.float() @ W.float().t()
    err = (out.float() - ref).abs().max().item()
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert err <= tol * max(1.0, ref.abs().max().item()), err


def test_linear_epilogues():
    torch.manual_seed(1)
    M, Nn, K = 300, 256, 512
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(Nn, K, device="cuda") * 0.1).half()
    bias = torch.randn(Nn, device="cuda")
    resid = torch.randn(M, Nn, device="cuda")
    rowmask = (torch.rand(M, device="cuda") > 0.3).float()
    ref_lin = A.float() @ W.float().t() + bias
    # bias + fp32 residual, in place
    out = resid.clone()
    _run(_desc(A, W, out, segs=[(0, 0, 0, K // 64)], R=M, bias=bias, resid=out, ldr=Nn))
    assert torch.allclose(out, ref_lin + resid, atol=5e-3, rtol=1e-3)
    # gelu(tanh) + pre-activation stash
    out = torch.empty(M, Nn, device="cuda", dtype=torch.half)
    pre = torch.empty(M, Nn, device="cuda", dtype=torch.half)
    _run(_desc(A, W, out, segs=[(0, 0, 0, K // 64)], R=M, bias=bias, act=N.ACT_GELU_TANH, aux_out=pre,
               ld_aux=Nn))
    assert torch.allclose(pre.float(), ref_lin, atol=1e-2, rtol=2e-3)
    assert torch.allclose(out.float(), torch.nn.functional.gelu(ref_lin, approximate="tanh"), atol=1e-2,
                          rtol=2e-3)
    # acc * gelu'(pre) (FF2 dgrad epilogue) with row mask
    out32 = torch.empty(M, Nn, device="cuda")
    _run(_desc(A, W, out32, segs=[(0, 0, 0, K // 64)], R=M, act=N.ACT_MUL_GELU_TANH_GRAD, mul_src=pre,
               ld_aux=Nn, rowmask=rowmask))
    p = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(p, approximate="tanh").sum().backward()
    ref = (A.float() @ W.float().t()) * p.grad * rowmask[:, None]
    assert torch.allclose(out32, ref, atol=5e-3, rtol=2e-3)
    # channel-major fp32 store with n_valid = 80 (final_proj)
    W80 = torch.zeros(128, K, device="cuda", dtype=torch.half)
    W80[:80] = W[:80]
    outT = torch.full((1, 80, M), float("nan"), device="cuda")
    _run(_desc(A, W80, outT, segs=[(0, 0, 0, K // 64)], R=M, bias=bias[:80].contiguous(), transposed_out=1,
               n_valid=80, rowmask=rowmask))
    ref = ((A.float() @ W.float().t()[:, :80] + bias[:80]) * rowmask[:, None]).t()[None]
    assert torch.allclose(outT, ref, atol=5e-3, rtol=1e-3)


@pytest.mark.parametrize("B,T,Cin", [(2, 200, 320), (3, 101, 256), (2, 128, 512), (1, 4, 256)])
def test_conv_k3(B, T, Cin):
    """Conv1d(k=3, padding=1) as three row-shifted segments on a token-major [B, T, Cin] tensor."""
    torch.manual_seed(2)
    Cout = 256
    x = (torch.randn(B, T, Cin, device="cuda") * 0.5).half()
    conv = torch.nn.Conv1d(Cin, Cout, 3, padding=1).cuda()
    w = conv.weight.detach().half()  # [Cout, Cin, 3]
    W2 = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()  # [n][tap*Cin + c]
    out = torch.full((B, T, Cout), float("nan"), device="cuda", dtype=torch.half)
    nk = Cin // 64
    d = _desc(x, W2, out, segs=[(0, -1, 0, nk), (0, 0, 0, nk), (0, 1, 0, nk)], R=T, nbatch=B,
              bias=conv.bias.detach().float().contiguous())
    _run(d)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), conv.bias.detach(), padding=1)
    ref = ref.transpose(1, 2)
    assert torch.allclose(out.float(), ref, atol=1e-2, rtol=3e-3), (out.float() - ref).abs().max()


def test_two_sources_and_interleaved_rows():
    """Channel-concat of two sources + output rows 2i+1 (ConvTranspose phase / strided dgrad)."""
    torch.manual_seed(3)
    B, T = 2, 77
    a0 = (torch.randn(B, T, 256, device="cuda") * 0.5).half()
    a1 = (torch.randn(B, T, 256, device="cuda") * 0.5).half()
    W = (torch.randn(256, 512, device="cuda") * 0.1).half()
    out = torch.zeros(B, 2 * T, 512, device="cuda", dtype=torch.half)
    d = _desc([a0, a1], W, out, segs=[(0, 0, 0, 4), (1, -1, 0, 4)], R=T, nbatch=B, rmul=2, roff=1,
              out_rows=2 * T, ldc=512, col_off=256, n_valid=256)
    _run(d)
    a1s = torch.zeros_like(a1)
    a1s[:, 1:] = a1[:, :-1]
    ref = torch.cat([a0, a1s], -1).float() @ W.float().t()
    assert torch.allclose(out[:, 1::2, 256:].float(), ref, atol=1e-2, rtol=3e-3)
    assert out[:, 0::2].abs().max().item() == 0 and out[:, :, :256].abs().max().item() == 0


def test_cluster_multicast_path_matches():
    """The opt-in A-tile multicast (thread-block clusters along N, CVFLOW_GEMM_CLUSTER=4) is read once per process,
    so the same parity tests run in a child process with it enabled."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CVFLOW_GEMM_CLUSTER="4")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", "not cluster_multicast"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,L", [(3, 77), (2, 400), (1, 1500), (5, 128)])
def test_conv_k3_groupnorm_partials(B, L, dtype):
    """GroupNorm statistics taken in the conv GEMM's epilogue (gn_part): Chan-merging the {n, mean, M2} partials of a
    sample gives the biased mean / variance of the fp32 conv output over ALL L rows x 32 channels of each group
    (what nn.GroupNorm(8, 256) reduces over, modules.py:60-73, padded positions included)."""
    torch.manual_seed(1)
    C_in, C_out = 256, 256
    x = (torch.randn(B, L, C_in, device="cuda") * 0.5).to(dtype)
    w = (torch.randn(C_out, 3 * C_in, device="cuda") * 0.05).to(dtype)
    bias = torch.randn(C_out, device="cuda") * 2.0 + 3.0          # a large mean stresses the one-pass partials
    out = torch.empty(B, L, C_out, device="cuda", dtype=dtype)
    nsplit = 4 * ((L + 127) // 128)
    part = torch.full((B, nsplit, C_out // 32, 3), float("nan"), device="cuda")
    d = _desc(x, w, out, segs=[(0, t - 1, 0, C_in // 64) for t in range(3)], R=L, nbatch=B, dtype=dtype, bias=bias,
              gn_part=part)
    _run(d)
    xp = torch.nn.functional.pad(x.float(), (0, 0, 1, 1))
    cols = torch.cat([xp[:, t:t + L] for t in range(3)], dim=2)
    ref = cols @ w.float().t() + bias
    assert torch.allclose(out.float(), ref, atol=3e-2, rtol=2e-2)
    assert torch.isfinite(part).all()
    n = part[..., 0].double()
    assert torch.equal(n.sum(1), torch.full((B, 8), 32.0 * L, device="cuda", dtype=torch.float64))
    mean = (n * part[..., 1].double()).sum(1) / n.sum(1)
    m2 = part[..., 2].double().sum(1) + (n * (part[..., 1].double() - mean[:, None]) ** 2).sum(1)
    var = m2 / n.sum(1)
    g = ref.double().view(B, L, 8, 32)
    assert torch.allclose(mean, g.mean(dim=(1, 3)), atol=2e-3, rtol=1e-3)
    assert torch.allclose(var, g.var(dim=(1, 3), unbiased=False), rtol=3e-3)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,K", [(6400, 512), (333, 1024), (12800, 512), (128, 64)])
def test_linear_fused_layernorm(M, K, dtype):
    """h = resid + A W^T + bias (fp32) and x~ = LayerNorm(h) (16-bit) from ONE launch: the 128x256 tile owns whole rows, the
    statistics are taken two-pass over the row kept in TMEM (replaces `h = h + to_out(o); x = norm3(h)`, modules.py:349-375)."""
    torch.manual_seed(2)
    A = (torch.randn(M, K, device="cuda") * 0.5).to(dtype)
    W = (torch.randn(256, K, device="cuda") * 0.1).to(dtype)
    bias = torch.randn(256, device="cuda")
    resid = torch.randn(M, 256, device="cuda") * 3.0 + 1.5        # a row mean away from zero
    gamma, beta = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda")
    out = resid.clone()
    xn = torch.full((M, 256), float("nan"), device="cuda").to(dtype)
    d = _desc(A, W, out, segs=[(0, 0, 0, K // 64)], R=M, dtype=dtype, bias=bias, resid=out, ldr=256, ln_gamma=gamma,
              ln_beta=beta, aux_out=xn, ld_aux=256)
    _run(d)
    ref_h = resid + A.float() @ W.float().t() + bias
    ref_x = torch.nn.functional.layer_norm(ref_h, (256,), gamma, beta, eps=1e-5)
    assert torch.allclose(out, ref_h, atol=2e-3, rtol=1e-3), float((out - ref_h).abs().max())
    tol = 4e-3 if dtype == torch.float16 else 3e-2
    assert torch.isfinite(xn.float()).all()
    assert torch.allclose(xn.float(), ref_x, atol=tol, rtol=tol), float((xn.float() - ref_x).abs().max())
