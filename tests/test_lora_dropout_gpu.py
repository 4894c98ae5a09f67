"""lora_dropout > 0 on the estimator's q/k/v LoRA branches (reference lora.py:66-74, the reference's default
configuration: config.py LORA_CONFIG / flow_lora use 0.05) through the C ABI.

The dropout draws cannot match torch's RNG stream, so parity is established with explicit masks: the golden
vector is the REAL reference with every LoRALinear's nn.Dropout replaced by a preset keep mask
(tests/golden/make_golden.py dropout); the CUDA path consumes the same masks through the debug-mask hook. The
production hash RNG is then tied to that verified path: a host replica of the hash produces the masks the device
must have used, and running with them explicitly has to reproduce the hash-mode result."""
import numpy as np
import pytest
import torch

from tests.helpers import build_estimator, dropout_masks, load_golden

pytestmark = pytest.mark.gpu


def _run(fx, est, mask=None, seed=None):
    from cosyvoice_lora_finetune_framework_b200 import _estimator as E
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    if seed is not None:
        torch.manual_seed(seed)
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    ne = E.native_of(est)
    if mask is not None:
        ne.set_debug_dropout_mask(mask.cuda())
    c = lambda k: fx[k].cuda()
    t = 1 - torch.cos(c("t_rand") * 0.5 * 3.14159265359)
    loss, _ = cfm._loss_with_noise(c("x1"), c("mask"), c("mu"), c("spks"), c("cond"), fx["prompt_lens"], t, c("z"),
                                   c("cfg_rand") > 0.2)
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), {k: p.grad.detach().cpu().clone() for k, p in est.named_parameters() if p.requires_grad}, ne


def _rel(grads, ref):
    num = sum((grads[k] - g).double().pow(2).sum() for k, g in ref.items())
    den = sum(g.double().pow(2).sum() for g in ref.values())
    return float((num / den).sqrt())


def test_lora_dropout_explicit_masks_vs_reference():
    dg = load_golden("dropout_tiny_prompt")
    fx = load_golden(dg["src"])
    est, _, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8, lora_dropout=dg["p"])
    est = est.cuda().train()
    keep = dropout_masks(dg["n_tbs"], dg["rows"], dg["p"], dg["mask_seed"])
    loss, grads, _ = _run(fx, est, mask=keep)
    assert abs(loss - float(dg["loss"])) <= 1e-2 * float(dg["loss"])
    assert set(grads) == set(dg["grads"])
    assert _rel(grads, dg["grads"]) <= 1e-2, _rel(grads, dg["grads"])
    assert _rel(grads, fx["grads"]) > 0.1            # and they are not the no-dropout gradients
    # eval(): dropout inactive, the folded path, the no-dropout golden
    est.zero_grad(set_to_none=True)
    est.eval()
    loss0, grads0, ne = _run(fx, est)
    assert ne._drop_active == 0.0
    assert abs(loss0 - float(fx["loss"])) <= 1e-2 * float(fx["loss"])
    assert _rel(grads0, fx["grads"]) <= 1e-2


def _host_keep(seed, n_tbs, rows, p):
    """Replica of lora_keep4() in csrc/lora_dropout.cu: one splitmix64 finaliser per (block, projection, token, feature
    quad) counter; its four 16-bit fields decide the quad's four features (dropped iff field < round(p 2^16))."""
    with np.errstate(over="ignore"):
        idx = np.arange(n_tbs * 3 * rows * 64, dtype=np.uint64)
        z = np.uint64(seed) + (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    thr = min(int(float(np.float32(p)) * 65536.0 + 0.5), 65535)
    fields = np.stack([(z >> np.uint64(16 * j)) & np.uint64(0xFFFF) for j in range(4)], axis=1)     # [quads][4]
    keep = (fields >= np.uint64(thr)).astype(np.uint8)
    return torch.from_numpy(keep.reshape(n_tbs, 3, rows, 256))


def test_lora_dropout_hash_rng_equals_its_explicit_masks():
    fx = load_golden("train_tiny_prompt")
    p = 0.1
    B, _, T = fx["x1"].shape
    n_tbs = 5
    est, _, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8, lora_dropout=p)
    loss_h, grads_h, ne = _run(fx, est.cuda().train(), seed=123)
    assert ne._drop_active == pytest.approx(p)
    seed_dev = (ne.dropout_seed() + 0x632BE59BD9B4E019) % (1 << 64)      # one bump before the first training forward
    keep = _host_keep(seed_dev, n_tbs, B * T, p)
    rate = float(keep.float().mean())
    assert abs(rate - (1 - p)) < 4 * (p * (1 - p) / keep.numel()) ** 0.5 + 1e-4, rate
    est2, _, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8, lora_dropout=p)
    loss_d, grads_d, _ = _run(fx, est2.cuda().train(), mask=keep, seed=123)
    assert abs(loss_h - loss_d) <= 1e-6 * abs(loss_d), (loss_h, loss_d)
    assert _rel(grads_h, grads_d) <= 1e-6, _rel(grads_h, grads_d)
    # a second step draws new masks (the device seed advances per training forward)
    est.zero_grad(set_to_none=True)
    loss_2, grads_2, _ = _run(fx, est)
    assert _rel(grads_2, grads_h) > 1e-3


def test_lora_dropout_training_step_r16_bf16():
    """The reference's flow_lora setting (r = 16, alpha = 32, dropout 0.05) in bf16: finite loss / gradients, fused
    optimiser step, and the un-folded operand images follow the updated LoRA parameters."""
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
    fx = load_golden("train_tiny")
    est, _, _ = build_estimator(1, 1, lora_r=16, lora_alpha=32, lora_dropout=0.05)
    est = est.cuda().train()
    est.cvflow_dtype = torch.bfloat16
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    tr = FlowLoRATrainer(cfm, lr=1e-3)
    c = lambda k: fx[k].cuda()
    torch.manual_seed(0)
    w0 = tr.ne.w0d.clone()
    losses = [float(tr.train_step(c("x1"), c("mask"), c("mu"), c("spks"), c("cond"))) for _ in range(3)]
    assert all(np.isfinite(losses)) and int(tr.found_inf.item()) == 0
    assert not torch.equal(tr.ne.w0d[:, :, 256:], w0[:, :, 256:])           # s B_cat refreshed after the step
    assert torch.equal(tr.ne.w0d[:, :, :256], w0[:, :, :256])              # frozen W0 untouched
    # the whole step, mask hash and operand refresh included, is CUDA-graph capturable; masks advance per replay
    g_losses = [float(tr.train_step_graphed(c("x1"), c("mask"), c("mu"), c("spks"), c("cond"))) for _ in range(3)]
    assert all(np.isfinite(g_losses)) and int(tr.found_inf.item()) == 0


def test_lora_dropout_300m_default_rate_vs_reference():
    """300M estimator at the reference's default lora_dropout = 0.05 (config.py:207-216), preset masks: loss, the stored
    gradients and every gradient norm of the real reference (fp16 operands: <= 1e-2)."""
    dg = load_golden("dropout_c1")
    fx = load_golden(dg["src"])
    est, _, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8, lora_dropout=dg["p"])
    est = est.cuda().train()
    keep = dropout_masks(dg["n_tbs"], dg["rows"], dg["p"], dg["mask_seed"])
    assert int(keep.sum()) == dg["keep_sum"]
    loss, grads, _ = _run(fx, est, mask=keep)
    assert abs(loss - float(dg["loss"])) <= 1e-2 * float(dg["loss"])
    assert _rel(grads, dg["grads"]) <= 1e-2, _rel(grads, dg["grads"])
    worst = max(abs(float(grads[k].norm()) - n) / (n + 1e-12) for k, n in dg["grad_norms"].items())
    assert worst <= 2e-2, worst


def test_lora_dropout_r16_explicit_masks_vs_oracle():
    """Rank 16 (the reference's flow_lora rank) exercises the 48-value reduction of the fused LayerNorm + masked
    down-projection kernel: CUDA path vs the CPU oracle on the same preset masks."""
    from oracle import flow_oracle as O
    from tests.helpers import attention_block_prefixes, lora_scaling_of, oracle_dropout_entries
    fx = load_golden("train_tiny_prompt")
    p = 0.2
    est, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=16, lora_alpha=32, lora_dropout=p)
    B, _, T = fx["x1"].shape
    prefixes = attention_block_prefixes(fx["n_blocks"], fx["n_mid"])
    keep = dropout_masks(len(prefixes), B * T, p, 77)
    P = {k: v.clone().requires_grad_(k.endswith(("lora_A", "lora_B"))) for k, v in sd.items()}
    P.update(oracle_dropout_entries(keep, p, prefixes, B, T))
    ref_loss, _, _ = O.cfm_compute_loss(P, fx["x1"], fx["mask"], fx["mu"], fx["spks"], fx["cond"], fx["prompt_lens"],
                                        fx["t_rand"], fx["z"], fx["cfg_rand"], lora_scaling=lora_scaling_of(sd, alpha=32))
    ref_loss.backward()
    ref = {k: P[k].grad for k in P if k.endswith(("lora_A", "lora_B"))}
    loss, grads, _ = _run(fx, est.cuda().train(), mask=keep)
    assert abs(loss - float(ref_loss)) <= 1e-2 * float(ref_loss)
    assert _rel(grads, ref) <= 1e-2, _rel(grads, ref)
