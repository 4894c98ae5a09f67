"""CPU-only tests: C ABI surface, LoRA semantics / checkpoint layout, module tree, schedules."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

from cosyvoice_lora_finetune_framework_b200 import _native, lora, modules, utils
from tests.helpers import GOLDEN, build_estimator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cvflow.h")).read()
    names = re.findall(r"CVFLOW_API[^;(]*?\b(cvflow_\w+)\s*\(", hdr)
    assert len(names) >= 18, names
    assert os.path.exists(_native.LIB_PATH), "build libcvflow.so first (python -m ...build)"
    L = ctypes.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libcvflow.so does not export %s" % n
    assert L.cvflow_abi_version() == 1
    L.cvflow_last_error.restype = ctypes.c_char_p
    assert isinstance(L.cvflow_last_error(), bytes)


def test_no_cpu_fallback():
    est, _, _ = build_estimator(1, 1)
    x = torch.zeros(1, 80, 8)
    with pytest.raises(RuntimeError):
        est(x, torch.ones(1, 1, 8), x, torch.zeros(1), torch.zeros(1, 80), x)
    with pytest.raises(RuntimeError):
        est.down_blocks[0][0](x, x, x)


def test_module_tree_matches_reference_key_layout():
    spec = torch.load(os.path.join(GOLDEN, "estimator_spec_300m.pt"), weights_only=False)
    est = modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0,
                                     attention_head_dim=64, n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn='gelu')
    sd = est.state_dict()
    assert list(sd.keys()) == spec["keys"]
    assert [tuple(v.shape) for v in sd.values()] == [tuple(s) for s in spec["shapes"]]
    assert sum(p.numel() for p in est.parameters()) == 81878864


def test_lora_injection_semantics():
    est = modules.ConditionalDecoder(320, 80, channels=(256, 256), dropout=0.0, attention_head_dim=64, n_blocks=4,
                                     num_mid_blocks=12, num_heads=8, act_fn='gelu')
    stats = lora.apply_lora_to_model(est, r=8, lora_alpha=16, lora_dropout=0.0,
                                     target_modules=['to_q', 'to_k', 'to_v', 'to_out'])
    # 'to_out' is a ModuleList whose Linear child is named "0": never wrapped (reference lora.py:178-182)
    assert stats["replaced_layers"] == 192 and stats["lora_params"] == 1179648
    assert stats["trainable_params"] == 1179648
    tb = est.mid_blocks[3][1][2]
    assert isinstance(tb.attn1.to_q, lora.LoRALinear) and isinstance(tb.attn1.to_out[0], nn.Linear)
    assert tb.attn1.to_q.scaling == 2.0
    assert float(tb.attn1.to_q.lora_B.abs().sum()) > 0           # B is N(0, 0.01), not zero
    assert all(('lora_' in n) == p.requires_grad for n, p in est.named_parameters())
    keys = list(est.state_dict().keys())
    assert "mid_blocks.3.1.2.attn1.to_q.original_layer.weight" in keys
    assert "mid_blocks.3.1.2.attn1.to_q.lora_A" in keys and "mid_blocks.3.1.2.attn1.to_q.lora_B" in keys


def test_lora_forward_merge_and_state_dict_layout():
    torch.manual_seed(0)
    net = nn.Sequential()
    net.add_module("to_q", nn.Linear(16, 8))
    net.add_module("other", nn.Linear(8, 4))
    net.add_module("to_v_conv", nn.Conv1d(4, 4, 1))
    net.register_buffer("stat", torch.ones(3))
    plain_keys = set(net.state_dict().keys())
    lora.apply_lora_to_model(net, r=2, lora_alpha=4, lora_dropout=0.0, target_modules=["to_q", "to_v"])
    x = torch.randn(5, 16)
    ll = net.to_q
    want = x @ ll.original_layer.weight.t() + ll.original_layer.bias + (x @ ll.lora_A.t() @ ll.lora_B.t()) * 2.0
    assert torch.allclose(ll(x), want, atol=1e-6)
    w0 = ll.original_layer.weight.clone()
    delta = ll.lora_B @ ll.lora_A * 2.0
    merged = lora.get_merged_state_dict(net)
    assert set(merged.keys()) == plain_keys
    # order: LoRA-wrapped layers first, then remaining params, then buffers (reference lora.py:300-320)
    assert list(merged.keys()) == ["to_q.weight", "to_q.bias", "to_v_conv.weight", "to_v_conv.bias", "other.weight",
                                   "other.bias", "stat"]
    assert torch.allclose(merged["to_q.weight"], w0 + delta, atol=1e-6)
    # merging mutates in place and is not idempotent (reference lora.py:264-279)
    again = lora.get_merged_state_dict(net)
    assert torch.allclose(again["to_q.weight"], w0 + 2 * delta, atol=1e-6)
    sd = lora.get_lora_state_dict(net)
    assert set(sd) == {"to_q.lora_A", "to_q.lora_B", "to_v_conv.lora_A.weight", "to_v_conv.lora_B.weight"}


def test_lora_save_load_roundtrip(tmp_path):
    a, _, _ = build_estimator(1, 1, lora_r=8)
    b, _, _ = build_estimator(1, 1, lora_r=8, wseed=77)
    path = str(tmp_path / "adapter.pt")
    lora.save_lora_weights(a, path)
    lora.load_lora_weights(b, path)
    for (ka, va), (kb, vb) in zip(lora.get_lora_state_dict(a).items(), lora.get_lora_state_dict(b).items()):
        assert ka == kb and torch.equal(va, vb)


def test_isolation_mask_and_schedule():
    m = modules.create_prompt_isolation_mask(5, 2, "cpu")
    assert m.shape == (1, 1, 5, 5)
    assert torch.isinf(m[0, 0, 3, 1]) and torch.isinf(m[0, 0, 0, 4]) and m[0, 0, 1, 0] == 0 and m[0, 0, 4, 2] == 0
    assert modules.create_prompt_isolation_mask(5, 0, "cpu").abs().sum() == 0
    assert modules.create_prompt_isolation_mask(5, 5, "cpu").abs().sum() == 0
    from cosyvoice_lora_finetune_framework_b200.trainer import lr_lambda
    assert lr_lambda(0, 50, 1000, 1e-4, 1e-6) == 0.0 and lr_lambda(25, 50, 1000, 1e-4, 1e-6) == 0.5
    assert abs(lr_lambda(50, 50, 1000, 1e-4, 1e-6) - 1.0) < 1e-9
    assert abs(lr_lambda(1000, 50, 1000, 1e-4, 1e-6) - 0.01) < 1e-6     # floor = min_lr / lr
    utils.set_all_random_seed(5)
    a = torch.rand(3)
    utils.set_all_random_seed(5)
    assert torch.equal(a, torch.rand(3))
    assert utils.pad_list([torch.ones(2), torch.ones(3)], 0).tolist() == [[1, 1, 0], [1, 1, 1]]


def test_cabi_reports_argument_errors_without_a_gpu():
    """Entry points validate their arguments before touching the device: int status + cvflow_last_error text, no
    exception, no exit (SURVEY section 8b 'Errors'). No compute is launched here."""
    from cosyvoice_lora_finetune_framework_b200 import _estimator as E
    L = E._lib()
    null = ctypes.c_void_p(None)
    cases = [
        (lambda: L.cvflow_estimator_forward(null, None, null), b"cvflow_estimator_forward"),
        (lambda: L.cvflow_estimator_backward(null, null, 1.0, null, null), b"cvflow_estimator_backward"),
        (lambda: L.cvflow_estimator_backward_inputs(null, null, 1.0, null, None, null), b"cvflow_estimator_backward_inputs"),
        (lambda: L.cvflow_set_lora_dropout(null, 0.1, ctypes.c_uint64(1), null, 0), b"cvflow_set_lora_dropout"),
        (lambda: L.cvflow_lora_refresh(null, null), b"cvflow_lora_refresh"),
        (lambda: L.cvflow_set_workspace(null, null, 0), b"cvflow_set_workspace"),
        (lambda: L.cvflow_cfm_prep(null, null, null, null, 1, 8, 1e-6, null), b"cvflow_cfm_prep"),
    ]
    for call, tag in cases:
        rc = call()
        assert rc < 0, tag
        assert tag in L.cvflow_last_error(), (tag, L.cvflow_last_error())
    assert L.cvflow_workspace_bytes(null, 2, 16, 1) < 0


def test_cabi_path_input_entry_points_validate_arguments():
    """The stateless path-input entry points (length regulator, pack, speaker affine) reject null / inconsistent
    arguments through the status code; the segment table must tile the output frames in order."""
    from cosyvoice_lora_finetune_framework_b200 import _path_inputs as PI
    L = PI._lib()
    null = ctypes.c_void_p(None)
    assert L.cvflow_regulator_saved_floats(2, 130) == 4 * 2 * 130 * 80 + 4 * 2 * 2 * 3
    assert L.cvflow_regulator_scratch_floats(2, 130) == 2 * 2 * 130 * 80 + 2 * 2 * 2 * 2
    assert L.cvflow_regulator_forward(None, None, null) < 0 and b"cvflow_regulator_forward" in L.cvflow_last_error()
    w, io = PI.RegulatorWeights(), PI.RegulatorIO()
    fake = 0x1000                                  # never dereferenced: validation fails first
    for i in range(5):
        w.wf[i] = w.wb[i] = w.bias[i] = fake
    for i in range(4):
        w.gamma[i] = w.beta[i] = fake
    io.src, io.saved, io.out = fake, fake, fake
    io.B, io.n_src, io.T, io.n_seg = 1, 10, 20, 2
    io.seg[0][:] = [0, 4, 0, 8]
    io.seg[1][:] = [4, 6, 9, 11]                   # gap: frame 8 is covered by no segment
    assert L.cvflow_regulator_forward(ctypes.byref(w), ctypes.byref(io), null) < 0
    assert b"segment 1" in L.cvflow_last_error()
    io.seg[1][:] = [4, 7, 8, 12]                   # reads source frame 10 of 10
    assert L.cvflow_regulator_forward(ctypes.byref(w), ctypes.byref(io), null) < 0
    io.seg[1][:] = [4, 6, 8, 11]                   # covers 19 of 20 frames
    assert L.cvflow_regulator_forward(ctypes.byref(w), ctypes.byref(io), null) < 0
    assert b"cover 19 frames" in L.cvflow_last_error()
    w.gamma[2] = None
    io.seg[1][:] = [4, 6, 8, 12]
    assert L.cvflow_regulator_backward(ctypes.byref(w), ctypes.byref(io), null, null, null, null) < 0
    assert b"weight image 2" in L.cvflow_last_error()
    assert L.cvflow_path_inputs_pack(null, null, 0, null, -6.0, 2.0, 0.0, null, null, null, 1, 8, null) < 0
    assert L.cvflow_spk_affine(null, null, null, null, 1, 192, 80, null) < 0 and b"cvflow_spk_affine" in L.cvflow_last_error()
    # the segment table of InterpolateRegulator.inference (modules.py:826-836)
    assert PI.inference_segments(30, 75, 52, 129) == [(0, 30, 0, 52), (30, 20, 52, 34), (50, 35, 86, 61), (85, 20, 147, 34)]
    assert PI.inference_segments(0, 33, 0, 56) == [(0, 33, 0, 56)]
