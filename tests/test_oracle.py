"""Pin the CPU oracle (oracle/flow_oracle.py) against vectors produced by the real reference
(tests/golden/make_golden.py) and against the reference's docstring known-answers."""
import pytest
import torch

from oracle import flow_oracle as O
from tests.helpers import build_estimator, load_golden, lora_scaling_of, wsum


def test_make_pad_mask_docstring_kat():
    # reference utils.py:28-33
    got = O.make_pad_mask(torch.tensor([5, 3, 2])).int().tolist()
    assert got == [[0, 0, 0, 0, 0], [0, 0, 0, 1, 1], [0, 0, 1, 1, 1]]
    from cosyvoice_lora_finetune_framework_b200 import utils
    assert utils.make_pad_mask(torch.tensor([5, 3, 2])).int().tolist() == got
    assert utils.make_pad_mask(torch.tensor([2, 1]), 4).int().tolist() == [[0, 0, 1, 1], [0, 1, 1, 1]]
    b = utils.mask_to_bias(torch.tensor([True, False]), torch.float32)
    assert b.tolist() == [0.0, -1.0e10]


@pytest.mark.parametrize("name", ["train_tiny", "train_tiny_prompt", "train_tiny_padprompt", "train_c1", "train_c1_prompt"])
def test_train_step_matches_reference(name):
    fx = load_golden(name)
    est, sd, stats = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    assert abs(wsum(sd) - fx["wsum"]) <= 1e-6 * fx["wsum"]
    assert stats["replaced_layers"] == fx["lora_stats"]["replaced_layers"]
    assert stats["lora_params"] == fx["lora_stats"]["lora_params"]
    P = {k: v.clone().requires_grad_(k.endswith(("lora_A", "lora_B"))) for k, v in sd.items()}
    loss, y, pred = O.cfm_compute_loss(P, fx["x1"], fx["mask"], fx["mu"], fx["spks"], fx["cond"], fx["prompt_lens"],
                                       fx["t_rand"], fx["z"], fx["cfg_rand"], lora_scaling=lora_scaling_of(sd))
    assert torch.allclose(y, fx["y"], atol=1e-6)
    assert torch.allclose(pred, fx["pred"], atol=2e-4, rtol=1e-4)
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * abs(float(fx["loss"]))
    loss.backward()
    for k, g in fx["grads"].items():
        assert torch.allclose(P[k].grad, g, atol=1e-6 + 1e-3 * float(g.abs().max()), rtol=1e-3), k
    for k, n in fx["grad_norms"].items():
        assert abs(float(P[k].grad.norm()) - n) <= 2e-3 * n + 1e-9, k


@pytest.mark.parametrize("name", ["estimator_tiny", "estimator_300m"])
def test_estimator_matches_reference(name):
    fx = load_golden(name)
    _, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"])
    assert abs(wsum(sd) - fx["wsum"]) <= 1e-6 * fx["wsum"]
    for c in fx["cases"]:
        with torch.no_grad():
            out = O.estimator_forward(sd, c["x"], c["mask"], c["mu"], c["t"], c["spks"], c["cond"])
        # the reference's own export check uses rtol 1e-2 / atol 1e-4 (export_onnx.py:115)
        assert torch.allclose(out, c["out"], rtol=1e-3, atol=1e-4), c["T"]


@pytest.mark.parametrize("name", ["euler_tiny", "euler_300m"])
def test_euler_matches_reference(name):
    fx = load_golden(name)
    _, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"])
    with torch.no_grad():
        mel, cache = O.cfm_forward(sd, fx["mu"], fx["mask"], fx["n_steps"], fx["z"].clone(), fx["spks"], fx["cond"],
                                   prompt_len=fx["prompt"])
    assert cache.shape == fx["cache"].shape and torch.equal(cache, fx["cache"])
    assert torch.allclose(mel, fx["mel"], atol=2e-3, rtol=1e-3)


@pytest.mark.parametrize("name", ["inputgrads_tiny_prompt", "inputgrads_c1"])
def test_input_gradients_match_reference(name):
    """dL/dmu, dL/dspks, dL/dcond of compute_loss: what the modules upstream of the estimator train on."""
    ig = load_golden(name)
    fx = load_golden(ig["src"])
    _, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    mu, spks, cond = (fx[k].clone().requires_grad_(True) for k in ("mu", "spks", "cond"))
    loss, _, _ = O.cfm_compute_loss(sd, fx["x1"], fx["mask"], mu, spks, cond, fx["prompt_lens"], fx["t_rand"], fx["z"],
                                    fx["cfg_rand"], lora_scaling=lora_scaling_of(sd))
    loss.backward()
    for got, key in ((mu.grad, "dmu"), (spks.grad, "dspks"), (cond.grad, "dcond")):
        want = ig[key]
        assert torch.allclose(got, want, atol=1e-7 + 1e-3 * float(want.abs().max()), rtol=1e-3), key
    # masked frames and CFG-dropped samples get exactly zero
    pad = fx["mask"].expand_as(mu) == 0
    assert float(ig["dmu"][pad].abs().sum()) == 0.0 and float(mu.grad[pad].abs().sum()) == 0.0


def test_lora_dropout_matches_reference():
    """lora_dropout > 0 (lora.py:66-74) with the nn.Dropout draws replaced by preset masks in the reference run."""
    from tests.helpers import attention_block_prefixes, dropout_masks, oracle_dropout_entries
    dg = load_golden("dropout_tiny_prompt")
    fx = load_golden(dg["src"])
    _, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    B, _, T = fx["x1"].shape
    keep = dropout_masks(dg["n_tbs"], dg["rows"], dg["p"], dg["mask_seed"])
    assert int(keep.sum()) == dg["keep_sum"]
    P = {k: v.clone().requires_grad_(k.endswith(("lora_A", "lora_B"))) for k, v in sd.items()}
    P.update(oracle_dropout_entries(keep, dg["p"], attention_block_prefixes(fx["n_blocks"], fx["n_mid"]), B, T))
    loss, _, _ = O.cfm_compute_loss(P, fx["x1"], fx["mask"], fx["mu"], fx["spks"], fx["cond"], fx["prompt_lens"],
                                    fx["t_rand"], fx["z"], fx["cfg_rand"], lora_scaling=lora_scaling_of(sd))
    assert abs(float(loss) - float(dg["loss"])) <= 1e-5 * abs(float(dg["loss"]))
    loss.backward()
    for k, g in dg["grads"].items():
        assert torch.allclose(P[k].grad, g, atol=1e-6 + 1e-3 * float(g.abs().max()), rtol=1e-3), k
    # the masks matter: the no-dropout gradients of the same step are far away
    far = sum((fx["grads"][k] - g).double().pow(2).sum() for k, g in dg["grads"].items())
    den = sum(g.double().pow(2).sum() for g in dg["grads"].values())
    assert float((far / den).sqrt()) > 0.1


def test_padprompt_fixture_hits_the_padding():
    """train_tiny_padprompt: the boundary window (25 frames after the prompt, weight 5) of two samples runs past their
    length, so padded frames carry loss weight (reference flow_model.py:184-193 does not re-mask them)."""
    fx = load_golden("train_tiny_padprompt")
    w = O.cfm_loss_weights(fx["mask"], fx["prompt_lens"])
    pad = fx["mask"] == 0
    assert float(w[pad].sum()) > 0
    assert float(w[1, 0, 40:55].min()) == 5.0 and float(w[0, 0, 50:64].min()) == 5.0 and float(w[0, 0, :50].max()) == 0.0


def test_benchmarked_train_shape_matches_reference():
    """The oracle at the BENCHMARKED shape (300M estimator, 32 x 400 ragged, bench.py's batch): loss and the stored
    gradients of the real reference. This is also the workload `bench.py --impl reference` times."""
    from tests.helpers import bench_train_inputs
    fx = load_golden("train_c3")
    _, sd, stats = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    assert abs(wsum(sd) - fx["wsum"]) <= 1e-6 * fx["wsum"]
    x1, mask, mu, spks, cond, t_rand, z, cfg = bench_train_inputs(fx)
    P = {k: v.clone().requires_grad_(k.endswith(("lora_A", "lora_B"))) for k, v in sd.items()}
    loss, y, _ = O.cfm_compute_loss(P, x1, mask, mu, spks, cond, None, t_rand, z, cfg, lora_scaling=lora_scaling_of(sd))
    assert abs(float(loss) - float(fx["loss"])) <= 1e-5 * abs(float(fx["loss"]))
    loss.backward()
    for k, g in fx["grads"].items():
        assert torch.allclose(P[k].grad, g, atol=1e-6 + 1e-3 * float(g.abs().max()), rtol=1e-3), k
    worst = max(abs(float(P[k].grad.norm()) - n) / (n + 1e-12) for k, n in fx["grad_norms"].items())
    assert worst <= 2e-3, worst


def test_benchmarked_euler_shape_matches_reference():
    """configs[1]: 10 Euler steps + CFG at T = 700 with a 200-frame prompt."""
    from tests.helpers import bench_euler_inputs
    fx = load_golden("euler_c2")
    _, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"])
    mu, spks, cond, mask, z = bench_euler_inputs(fx)
    with torch.no_grad():
        mel, cache = O.cfm_forward(sd, mu, mask, fx["n_steps"], z.clone(), spks, cond, prompt_len=fx["prompt"])
    assert cache.shape == fx["cache"].shape and torch.equal(cache, fx["cache"])
    assert torch.allclose(mel, fx["mel"], atol=2e-3, rtol=1e-3)


def test_lora_dropout_300m_matches_reference():
    """300M-scale lora_dropout fixture at the reference's default p = 0.05 (config.py:207-216)."""
    from tests.helpers import attention_block_prefixes, dropout_masks, oracle_dropout_entries
    dg = load_golden("dropout_c1")
    fx = load_golden(dg["src"])
    _, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    B, _, T = fx["x1"].shape
    keep = dropout_masks(dg["n_tbs"], dg["rows"], dg["p"], dg["mask_seed"])
    assert int(keep.sum()) == dg["keep_sum"]
    P = {k: v.clone().requires_grad_(k.endswith(("lora_A", "lora_B"))) for k, v in sd.items()}
    P.update(oracle_dropout_entries(keep, dg["p"], attention_block_prefixes(fx["n_blocks"], fx["n_mid"]), B, T))
    loss, _, _ = O.cfm_compute_loss(P, fx["x1"], fx["mask"], fx["mu"], fx["spks"], fx["cond"], fx["prompt_lens"],
                                    fx["t_rand"], fx["z"], fx["cfg_rand"], lora_scaling=lora_scaling_of(sd))
    assert abs(float(loss) - float(dg["loss"])) <= 1e-5 * abs(float(dg["loss"]))
    loss.backward()
    for k, g in dg["grads"].items():
        assert torch.allclose(P[k].grad, g, atol=1e-6 + 1e-3 * float(g.abs().max()), rtol=1e-3), k
    worst = max(abs(float(P[k].grad.norm()) - n) / (n + 1e-12) for k, n in dg["grad_norms"].items())
    assert worst <= 2e-3, worst


# ------------------------------------------------------------------------------------------------------
# the inputs of the path: length regulator, speaker affine, conditioning pack
# ------------------------------------------------------------------------------------------------------
def test_interp_taps_bit_identical_to_torch():
    """north_star: bit-exact length-regulation indices. The oracle's taps reproduce F.interpolate(mode='linear')'s
    index arithmetic (what the reference calls, modules.py:822-836): the full weight matrix is read off torch with an
    identity probe and must be equal bit for bit, for up- and down-sampling, edge lengths and the benchmarked ratios."""
    import numpy as np
    import torch.nn.functional as F
    pairs = [(n_in, n_out) for n_in in (1, 2, 3, 5, 20, 33, 35, 47, 75, 100, 232, 233, 406, 870)
             for n_out in (1, 2, 7, 34, 56, 61, 81, 129, 137, 200, 399, 400, 401, 700, 1500)]
    for n_in, n_out in pairs:
        i0, i1, w0, w1 = O.interp_linear_taps(n_in, n_out)
        want = F.interpolate(torch.eye(n_in)[None], size=n_out, mode='linear')[0].numpy()
        got = np.zeros((n_in, n_out), dtype=np.float32)
        for j in range(n_out):
            got[i0[j], j] += w0[j]
            got[i1[j], j] += w1[j]
        assert np.array_equal(got, want), (n_in, n_out)


def test_regulator_oracle_matches_reference():
    from tests.helpers import close_sums, regulator_inputs
    fx = load_golden("regulator_tiny")
    sd, x, yl, R = regulator_inputs(fx)
    x.requires_grad_(True)
    out = O.regulator_forward(sd, "", x, yl)
    (out * R).sum().backward()
    assert torch.allclose(out, fx["out"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(x.grad, fx["dx"], atol=1e-5, rtol=1e-4)
    assert float(out[1, 60:].abs().max()) == 0.0 and float(out[2, 33:].abs().max()) == 0.0      # padded frames are exact zeros
    with torch.no_grad():
        o1 = O.regulator_inference(sd, "", fx["inf_x1"], fx["inf_x2"], *fx["inf_len"])
        o2 = O.regulator_inference(sd, "", fx["inf_short_x"][:, :0], fx["inf_short_x"], 0, fx["inf_short_len"])
    assert o1.shape[1] == fx["inf_total"] and torch.allclose(o1, fx["inf_out"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(o2, fx["inf_short_out"], atol=1e-5, rtol=1e-5)
    # the benchmarked shape (32 utterances, 232 tokens -> 400 ragged frames): checksums and sampled rows of the reference
    fc = load_golden("regulator_c3")
    sd, x, yl, R = regulator_inputs(fc)
    x.requires_grad_(True)
    out = O.regulator_forward(sd, "", x, yl)
    (out * R).sum().backward()
    assert close_sums(out.detach(), fc["out_sum"], 1e-5) and close_sums(x.grad, fc["dx_sum"], 1e-5)
    assert torch.allclose(out.detach()[:, ::97], fc["out_rows"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(x.grad[:, ::53], fc["dx_rows"], atol=1e-5, rtol=1e-4)


def test_path_inputs_oracle_vs_reference_flow_model():
    """Speaker affine and the conditioning pack against what the REAL MaskedDiffWithXvec.forward handed to compute_loss
    (fixture flowmodel_tiny: x1, cond, spks seen by the reference)."""
    from cosyvoice_lora_finetune_framework_b200 import flow_model, utils
    fx = load_golden("flowmodel_tiny")
    utils.set_all_random_seed(fx["init_seed"])
    m = flow_model.build_flow_model(None, 'cpu', **fx["arch"])
    b = fx["batch"]
    spks = O.speaker_affine(m.spk_embed_affine_layer.weight, m.spk_embed_affine_layer.bias, b["embedding"])
    assert torch.allclose(spks, fx["spks"], atol=1e-6, rtol=1e-5)
    desc = [(int(n), int(p), 0, int(c) > 0) for n, p, c in zip(b["speech_feat_len"], fx["prompt_lens"], b["cross_sample_mel_len"])]
    assert any(d[3] for d in desc)                      # the fixture exercises the cross-sample prompt (strategy 5)
    x1, cond, mask = O.path_inputs_pack(b["speech_feat"], b["cross_sample_mel"], desc, m.mel_mean, m.mel_std, 0.0)
    assert torch.allclose(x1, fx["x1"], atol=1e-6, rtol=1e-6) and torch.allclose(cond, fx["cond"], atol=1e-6, rtol=1e-6)
