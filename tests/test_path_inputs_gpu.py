"""The inputs of the path through the C ABI: the length regulator (csrc/regulator.cu) with its autograd, the
speaker affine layer and the conditioning pack, against the REAL reference (tests/golden/regulator_*.pt from
InterpolateRegulator, modules.py:800-837) and the CPU oracle.

Tolerances: interpolation indices and weights bit-exact (north_star); the fp32 conv stack 1e-4 (different summation order
only); integer / layout work (mask, zeros, transposes) exact."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import flow_oracle as O
from tests.helpers import build_regulator, close_sums, load_golden, regulator_inputs

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def test_interpolation_taps_bit_exact():
    """The regulator's first launch interpolates in its prologue. With an identity first convolution (centre tap 1,
    bias 0) its raw output IS the interpolated operand, and with a one-hot probe (source frame s lights channel s % 80 of
    utterance s // 80) every product is exact: the kernel's tap matrix must equal the oracle's (== torch's) bit for bit."""
    from cosyvoice_lora_finetune_framework_b200 import _native as N, _path_inputs as PI
    L = PI._lib()
    sd0, _, _, _ = regulator_inputs(load_golden("regulator_tiny"))
    sd = {k: v.clone() for k, v in sd0.items()}
    sd["model.0.weight"].zero_()
    sd["model.0.weight"][torch.arange(80), torch.arange(80), 1] = 1.0
    sd["model.0.bias"].zero_()
    reg = build_regulator(sd).cuda()
    im = PI.images_of(reg)
    for n_src, T in ((47, 81), (232, 400), (233, 399), (5, 320), (300, 7), (1, 9), (240, 1500)):
        B = (n_src + 79) // 80
        x = torch.zeros(B, n_src, 80)
        s = torch.arange(n_src)
        x[s // 80, s, s % 80] = 1.0
        xd = x.cuda()
        saved = torch.empty(L.cvflow_regulator_saved_floats(B, T), device="cuda")
        out = torch.empty(B, T, 80, device="cuda")
        io = PI._io(xd, T, ((0, n_src, 0, T),), None, None, out, False, saved)
        N.check(L.cvflow_regulator_forward(C.byref(im.c), C.byref(io), PI._stream()))
        torch.cuda.synchronize()
        y0 = saved[: B * T * 80].view(B, T, 80).cpu()
        i0, i1, w0, w1 = O.interp_linear_taps(n_src, T)
        want = np.zeros((n_src, T), dtype=np.float32)
        for j in range(T):
            want[i0[j], j] += w0[j]
            want[i1[j], j] += w1[j]
        got = np.zeros_like(want)
        for si in range(n_src):
            got[si] = y0[si // 80, :, si % 80].numpy()
        assert np.array_equal(got, want), (n_src, T)


def test_regulator_forward_backward_vs_reference():
    fx = load_golden("regulator_tiny")
    sd, x, yl, R = regulator_inputs(fx)
    reg = build_regulator(sd).cuda()
    xd = x.cuda().requires_grad_(True)
    out, yl2 = reg(xd, yl)
    (out * R.cuda()).sum().backward()
    assert yl2 is yl and out.shape == fx["out"].shape
    assert torch.allclose(out.detach().cpu(), fx["out"], atol=1e-4, rtol=1e-4), float((out.detach().cpu() - fx["out"]).abs().max())
    assert _rel(xd.grad.cpu(), fx["dx"]) <= 1e-4, _rel(xd.grad.cpu(), fx["dx"])
    o = out.detach().cpu()
    assert float(o[1, 60:].abs().max()) == 0.0 and float(o[2, 33:].abs().max()) == 0.0       # padded frames: exact zeros
    print("regulator_tiny: out max-abs %.2e, dx rel-L2 %.2e" % (float((o - fx["out"]).abs().max()), _rel(xd.grad.cpu(), fx["dx"])))


def test_regulator_inference_layouts_vs_reference():
    """InterpolateRegulator.inference: prompt | head 20 | middle | tail 20 stretched separately (modules.py:826-836),
    and the short-target branch without a prompt."""
    fx = load_golden("regulator_tiny")
    sd, _, _, _ = regulator_inputs(fx)
    reg = build_regulator(sd).cuda()
    with torch.no_grad():
        o1, n1 = reg.inference(fx["inf_x1"].cuda(), fx["inf_x2"].cuda(), *fx["inf_len"])
        xs = fx["inf_short_x"].cuda()
        o2, n2 = reg.inference(xs[:, :0], xs, 0, fx["inf_short_len"])
    assert n1 == fx["inf_total"] and n2 == fx["inf_short_len"]
    assert torch.allclose(o1.cpu(), fx["inf_out"], atol=1e-4, rtol=1e-4)
    assert torch.allclose(o2.cpu(), fx["inf_short_out"], atol=1e-4, rtol=1e-4)


def test_regulator_benchmarked_shape_vs_reference():
    """32 utterances, 232 tokens -> 400 ragged frames (the shape bench.py's flow_model_leg runs): checksums and sampled
    rows of the real reference's output and input gradient."""
    fx = load_golden("regulator_c3")
    sd, x, yl, R = regulator_inputs(fx)
    reg = build_regulator(sd).cuda()
    xd = x.cuda().requires_grad_(True)
    out, _ = reg(xd, yl)
    (out * R.cuda()).sum().backward()
    o, g = out.detach().cpu(), xd.grad.cpu()
    assert close_sums(o, fx["out_sum"], 1e-5) and close_sums(g, fx["dx_sum"], 1e-5)
    assert torch.allclose(o[:, ::97], fx["out_rows"], atol=1e-4, rtol=1e-4)
    assert _rel(g[:, ::53], fx["dx_rows"]) <= 1e-4
    print("regulator_c3: rows max-abs %.2e, dx rows rel-L2 %.2e" % (float((o[:, ::97] - fx["out_rows"]).abs().max()),
                                                                   _rel(g[:, ::53], fx["dx_rows"])))


def test_regulator_blinding_and_channel_major_vs_oracle():
    """The layout compute_loss takes ([B][80][T]) with the text-side blinding (flow_model.py:372-373) folded in:
    forward and input gradient against the oracle followed by the reference's zeroing and transpose."""
    from cosyvoice_lora_finetune_framework_b200 import _path_inputs as PI
    fx = load_golden("regulator_tiny")
    sd, x, yl, R = regulator_inputs(fx)
    reg = build_regulator(sd).cuda()
    blind = [9, 0, 5]
    xr = x.clone().requires_grad_(True)
    ref = O.regulator_forward(sd, "", xr, yl).clone()
    for i, p in enumerate(blind):
        ref[i, :p] = 0.0
    ref = ref.transpose(1, 2)
    Rt = R.transpose(1, 2).contiguous()
    (ref * Rt).sum().backward()
    xd = x.cuda().requires_grad_(True)
    out = PI.regulate(reg, xd, int(yl.max()), lens=yl, blind=blind, channel_major=True)
    (out * Rt.cuda()).sum().backward()
    assert out.shape == ref.shape and torch.allclose(out.detach().cpu(), ref.detach(), atol=1e-4, rtol=1e-4)
    assert float(out[0, :, :9].abs().max()) == 0.0 and float(out[2, :, :5].abs().max()) == 0.0
    assert _rel(xd.grad.cpu(), xr.grad) <= 1e-4


def test_regulator_refuses_trainable_parameters():
    fx = load_golden("regulator_tiny")
    sd, x, yl, _ = regulator_inputs(fx)
    reg = build_regulator(sd).cuda()
    next(reg.parameters()).requires_grad_(True)
    with pytest.raises(RuntimeError, match="frozen"):
        reg(x.cuda(), yl)


def test_path_inputs_pack_and_speaker_affine_vs_oracle():
    from cosyvoice_lora_finetune_framework_b200 import _path_inputs as PI
    g = torch.Generator().manual_seed(3)
    B, T, Tc = 4, 70, 33
    feat = torch.randn(B, T, 80, generator=g) * 2 - 6
    cross = torch.randn(B, Tc, 80, generator=g) * 2 - 6
    desc = [(70, 12, 0, 0), (41, 9, 5, 1), (55, 0, 0, 0), (33, 20, 13, 1)]      # (len, prompt, gap, from cross)
    mean, std, sil = -6.0, 2.0, (-11.5 + 6.0) / 2.0
    rx1, rcond, rmask = O.path_inputs_pack(feat, cross, desc, mean, std, sil)
    x1, cond, mask = PI.pack_inputs(feat.cuda(), cross.cuda(), [list(map(int, d)) for d in desc], mean, std, sil)
    assert torch.equal(x1.cpu(), rx1) and torch.equal(cond.cpu(), rcond) and torch.equal(mask.cpu(), rmask)
    lin = torch.nn.Linear(192, 80)
    for p in lin.parameters():
        p.requires_grad_(False)
    e = torch.randn(5, 192, generator=g)
    ref = O.speaker_affine(lin.weight, lin.bias, e)
    got = PI.spk_affine(lin.cuda(), e.cuda())
    assert torch.allclose(got.cpu(), ref, atol=1e-6, rtol=1e-5)
