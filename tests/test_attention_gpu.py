"""attn1 kernels on their own against a plain fp32 PyTorch restatement of Attention.forward
(reference modules.py:253-293 with the additive bias of utils.py:103-109 and the prompt-isolation
mask of modules.py:844-879), forward and backward, through the C ABI."""
import pytest
import torch

from cosyvoice_lora_finetune_framework_b200 import _estimator as E
from cosyvoice_lora_finetune_framework_b200 import _native as N

pytestmark = pytest.mark.gpu


def _reference(qkv, mask, iso_p, dout=None):
    """fp32 attention on 16-bit-rounded inputs. qkv [B,L,1536], mask [B,L] -> o [B,L,512] (+ grads)."""
    B, L, _ = qkv.shape
    x = qkv.float().detach().requires_grad_(dout is not None)
    q, k, v = [t.view(B, L, 8, 64).transpose(1, 2) for t in x.split(512, dim=-1)]
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * 0.125
    bias = (1.0 - mask[:, None, None, :].expand(B, 1, L, L)) * -1e10
    if 0 < iso_p < L:
        idx = torch.arange(L, device=qkv.device)
        cross = (idx[:, None] < iso_p) != (idx[None, :] < iso_p)
        bias = bias.masked_fill(cross[None, None], float("-inf"))
    attn = (sim + bias).softmax(dim=-1)
    o = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(B, L, 512)
    if dout is None:
        return o, None
    o.backward(dout.float())
    return o.detach(), x.grad


def _run_fwd(qkv, mask, iso_p, dt):
    L_ = E._lib()
    B, L, ldq = qkv.shape
    o = torch.full((B, L, 512), float("nan"), device="cuda", dtype=dt)
    lse = torch.full((B, 8, L), float("nan"), device="cuda")
    kmax = torch.zeros(E._lib().cvflow_attention_scratch_ints(B, L), dtype=torch.int32, device="cuda")
    N.check(L_.cvflow_attention_forward(qkv.data_ptr(), ldq, B, L, N.dtype_code(dt), mask.data_ptr(), kmax.data_ptr(), iso_p,
                                        o.data_ptr(), lse.data_ptr(), E._stream()))
    torch.cuda.synchronize()
    return o, lse, kmax


CASES = [
    # B, L, lengths, iso_p
    (2, 48, [48, 31], 0),
    (3, 100, [100, 61, 17], 0),
    (4, 200, [200, 160, 128, 121], 0),
    (2, 200, [200, 150], 37),
    (3, 400, [400, 257, 250], 0),
    (2, 350, [350, 300], 100),
    (2, 700, [700, 512], 200),
    (2, 750, [750, 333], 0),
    (1, 1500, [1500], 0),
]


@pytest.mark.parametrize("B,L,lens,iso_p", CASES)
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_attention_forward(B, L, lens, iso_p, dt):
    torch.manual_seed(L + B)
    qkv = (torch.randn(B, L, 1536, device="cuda") * 1.5).to(dt)
    mask = (torch.arange(L, device="cuda")[None, :] < torch.tensor(lens, device="cuda")[:, None]).float()
    o, lse, kmax = _run_fwd(qkv, mask, iso_p, dt)
    assert kmax[:B].tolist() == lens
    ref, _ = _reference(qkv, mask, iso_p)
    tol = 4e-3 if dt == torch.float16 else 2e-2
    for b in range(B):
        n = lens[b]
        got = o[b, :n].float()
        assert torch.isfinite(got).all()
        err = (got - ref[b, :n]).abs().max().item()
        assert err <= tol * max(1.0, ref[b, :n].abs().max().item()), (b, err)
        # rows of pure padding beyond the last 128-row tile that holds a valid row are defined zeros
        first_skipped = ((n + 127) // 128) * 128
        if first_skipped < L:
            assert o[b, first_skipped:].float().abs().max().item() == 0.0
            assert torch.isinf(lse[b, :, first_skipped:]).all()
    assert torch.isfinite(o.float()).all()


@pytest.mark.parametrize("B,L,lens,iso_p", CASES)
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_attention_backward(B, L, lens, iso_p, dt):
    L_ = E._lib()
    torch.manual_seed(7 * L + B)
    qkv = (torch.randn(B, L, 1536, device="cuda") * 1.2).to(dt)
    lens_t = torch.tensor(lens, device="cuda")
    mask = (torch.arange(L, device="cuda")[None, :] < lens_t[:, None]).float()
    # the gradient of padded rows is zero in the estimator (every consumer masks them)
    dout = (torch.randn(B, L, 512, device="cuda") * mask[:, :, None]).to(dt)
    o, lse, kmax = _run_fwd(qkv, mask, iso_p, dt)
    delta = torch.zeros(B, 8, L, device="cuda")
    dqkv = torch.full((B, L, 1536), float("nan"), device="cuda", dtype=dt)
    N.check(L_.cvflow_attention_backward(qkv.data_ptr(), 1536, B, L, N.dtype_code(dt), mask.data_ptr(), kmax.data_ptr(), iso_p,
                                         o.data_ptr(), lse.data_ptr(), dout.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                         E._stream()))
    torch.cuda.synchronize()
    _, gref = _reference(qkv, mask, iso_p, dout)
    assert torch.isfinite(dqkv.float()).all()
    tol = 6e-3 if dt == torch.float16 else 3e-2
    for b in range(B):
        n = lens[b]
        got, ref = dqkv[b, :n].float(), gref[b, :n]
        for name, sl in (("dq", slice(0, 512)), ("dk", slice(512, 1024)), ("dv", slice(1024, 1536))):
            err = (got[:, sl] - ref[:, sl]).abs().max().item()
            assert err <= tol * max(1.0, ref[:, sl].abs().max().item()), (b, name, err)
        # padded rows: dk = dv = 0 exactly (their probability is 0), dq = 0 because dout = 0 there
        if n < L:
            assert dqkv[b, n:].float().abs().max().item() == 0.0


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_attention_mask_with_holes(dt):
    """The key mask is an arbitrary {0,1} row, not only a prefix (utils.py:103-109 builds the bias from any mask): keys masked
    in the middle of a sample get probability 0, forward and backward; the valid extent is 1 + the last non-zero index."""
    L_ = E._lib()
    torch.manual_seed(11)
    B, L = 3, 300
    qkv = (torch.randn(B, L, 1536, device="cuda") * 1.3).to(dt)
    mask = torch.ones(B, L, device="cuda")
    mask[0, 40:75] = 0
    mask[1, 0:10] = 0
    mask[1, 130:190] = 0
    mask[1, 260:] = 0
    mask[2, ::3] = 0
    ext = [300, 260, 300]
    dout = (torch.randn(B, L, 512, device="cuda")).to(dt)
    for b in range(B):
        dout[b, ext[b]:] = 0
    o, lse, kmax = _run_fwd(qkv, mask, 0, dt)
    assert kmax[:B].tolist() == ext
    delta = torch.zeros(B, 8, L, device="cuda")
    dqkv = torch.full((B, L, 1536), float("nan"), device="cuda", dtype=dt)
    N.check(L_.cvflow_attention_backward(qkv.data_ptr(), 1536, B, L, N.dtype_code(dt), mask.data_ptr(), kmax.data_ptr(), 0,
                                         o.data_ptr(), lse.data_ptr(), dout.data_ptr(), delta.data_ptr(), dqkv.data_ptr(),
                                         E._stream()))
    torch.cuda.synchronize()
    ref, gref = _reference(qkv, mask, 0, dout)
    tol = 6e-3 if dt == torch.float16 else 3e-2
    for b in range(B):
        n = ext[b]
        assert (o[b, :n].float() - ref[b, :n]).abs().max().item() <= tol * max(1.0, ref[b, :n].abs().max().item())
        got, want = dqkv[b, :n].float(), gref[b, :n]
        assert (got - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
        dead = mask[b, :n] == 0   # masked keys receive no gradient through k and v
        assert dqkv[b, :n][dead][:, 512:].float().abs().max().item() == 0.0
