"""CFM training step on the GPU (fused prep -> estimator fwd -> loss -> bwd through the C ABI) vs the
reference's golden vectors (loss, every LoRA gradient norm, full gradients of selected layers) and
vs the CPU oracle. Tolerances are the north-star ones: loss and LoRA grads <= 1e-2 relative with
16-bit operands / fp32 accumulation (fp16 operands; the bf16 noise floor is reported separately,
BASELINE.md section 3)."""
import pytest
import torch

from tests.helpers import build_estimator, load_golden

pytestmark = pytest.mark.gpu


def _step(fx, dtype):
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    est, sd, stats = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    est = est.cuda().train()
    est.cvflow_dtype = dtype
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    c = lambda k: fx[k].cuda()
    t = 1 - torch.cos(c("t_rand") * 0.5 * 3.14159265359)
    keep = c("cfg_rand") > 0.2
    loss, y = cfm._loss_with_noise(c("x1"), c("mask"), c("mu"), c("spks"), c("cond"), fx["prompt_lens"], t, c("z"), keep)
    loss.backward()
    grads = {k: p.grad.detach().cpu().clone() for k, p in est.named_parameters() if p.requires_grad}
    return loss.item(), y.cpu(), grads, est


@pytest.mark.parametrize("name", ["train_tiny", "train_tiny_prompt", "train_tiny_padprompt", "train_c1", "train_c1_prompt"])
def test_train_step_fp16(name):
    fx = load_golden(name)
    loss, y, grads, est = _step(fx, torch.float16)
    ref_loss = float(fx["loss"])
    assert abs(loss - ref_loss) <= 1e-2 * abs(ref_loss), (loss, ref_loss)
    assert torch.allclose(y, fx["y"], atol=1e-5)
    assert set(grads) == set(fx["grad_norms"])
    # whole-bucket relative L2 error on the tensors stored in full
    num = sum((grads[k] - g).double().pow(2).sum() for k, g in fx["grads"].items())
    den = sum(g.double().pow(2).sum() for g in fx["grads"].values())
    rel = float((num / den).sqrt())
    assert rel <= 1e-2, rel
    # every tensor's norm
    worst = max(abs(float(grads[k].norm()) - n) / (n + 1e-12) for k, n in fx["grad_norms"].items())
    assert worst <= 2e-2, worst
    tot = float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values())))
    assert abs(tot - fx["grad_total_norm"]) <= 1e-2 * fx["grad_total_norm"]


def test_train_step_bf16_noise_floor():
    fx = load_golden("train_c1")
    loss, y, grads, _ = _step(fx, torch.bfloat16)
    assert abs(loss - float(fx["loss"])) <= 1e-2 * float(fx["loss"])
    num = sum((grads[k] - g).double().pow(2).sum() for k, g in fx["grads"].items())
    den = sum(g.double().pow(2).sum() for g in fx["grads"].values())
    # the reference's own bf16 autocast sits at 1.6e-2 rel-L2 vs its fp32 at this shape:
    # bf16 operands are held to 1.5x the reference's own floor
    rel = float((num / den).sqrt())
    print("train_c1 bf16: LoRA-grad bucket rel-L2 %.3e (reference under bf16 autocast: 1.6e-2)" % rel)
    assert rel <= 1.5 * 1.6e-2, rel


# ------------------------------------------------------------------------------------------------------
# The BENCHMARKED shape: 300M estimator, 32 x 400 ragged frames (bench.py's batch, seed 99), against the real
# reference's fp32 step (tests/golden/train_c3.pt; inputs regenerated from the seeds and checksum-verified).
# ------------------------------------------------------------------------------------------------------
def _bench_step(fx, dtype, graphed):
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    from tests.helpers import bench_train_inputs
    x1, mask, mu, spks, cond, t_rand, z, cfg = (v.cuda() for v in bench_train_inputs(fx))
    est, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    est = est.cuda().train()
    est.cvflow_dtype = dtype
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    t = 1 - torch.cos(t_rand * 0.5 * 3.14159265359)
    keep = cfg > 0.2

    def body():
        loss, _ = cfm._loss_with_noise(x1, mask, mu, spks, cond, None, t, z, keep)
        loss.backward()
        return loss.detach()

    if not graphed:
        loss = body()
    else:       # the same launches captured into one CUDA graph and replayed (the form bench.py times)
        from cosyvoice_lora_finetune_framework_b200 import _estimator as E
        ne = E.native_of(est)
        ne.attach_grads()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ne.grad_bucket.zero_()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = body()
        ne.grad_bucket.zero_()
        g.replay()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu().clone() for k, p in est.named_parameters() if p.requires_grad}
    return float(loss), grads


@pytest.mark.parametrize("graphed", [False, True])
def test_benchmarked_shape_fp16(graphed):
    from tests.helpers import grad_errors
    fx = load_golden("train_c3")
    loss, grads = _bench_step(fx, torch.float16, graphed)
    ref = float(fx["loss"])
    e = grad_errors(grads, fx["grads"])
    # per-tensor norms: tensors that carry >= 1 % of the largest gradient norm must be within 2e-2; the few tensors with
    # vanishing gradients (16-bit operands underflow there) are held to 1.5x what the REFERENCE ITSELF loses on its worst
    # tensor under fp16 autocast + loss scale at this shape (recorded in the fixture: 0.42)
    nmax = max(fx["grad_norms"].values())
    worst = max(abs(float(grads[k].norm()) - n) / (n + 1e-12) for k, n in fx["grad_norms"].items() if n >= 1e-2 * nmax)
    floor = fx["ref_autocast"]["fp16"]
    print("train_c3 fp16 graphed=%s: loss rel %.2e, %s, worst norm rel (significant tensors) %.2e; reference's own fp16 "
          "floor %s" % (graphed, abs(loss - ref) / ref, e, worst, floor))
    assert abs(loss - ref) <= 1e-2 * ref, (loss, ref)
    assert e["bucket_rel_l2"] <= 1e-2, e
    assert e["tensor_rel_l2_median"] <= 1e-2, e
    assert e["tensor_rel_l2_max"] <= 1.5 * floor["tensor_rel_l2_max"], (e, floor)
    assert worst <= 2e-2, worst
    tot = float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads.values())))
    assert abs(tot - fx["grad_total_norm"]) <= 1e-2 * fx["grad_total_norm"]


@pytest.mark.parametrize("graphed", [False, True])
def test_benchmarked_shape_bf16(graphed):
    """bf16 operands (the dtype bench.py runs): loss <= 1e-2; LoRA gradients within 1.5x the error the REFERENCE ITSELF
    makes under bf16 autocast at this exact shape (recorded in the fixture from the real reference: bucket rel-L2
    1.3e-2, per-tensor max 3.9e-2), because bf16's 8-bit mantissa puts the reference's own gradients above 1e-2."""
    from tests.helpers import grad_errors
    fx = load_golden("train_c3")
    floor = fx["ref_autocast"]["bf16"]
    loss, grads = _bench_step(fx, torch.bfloat16, graphed)
    ref = float(fx["loss"])
    e = grad_errors(grads, fx["grads"])
    print("train_c3 bf16 graphed=%s: loss rel %.2e, %s; reference's own bf16 floor %s" % (graphed, abs(loss - ref) / ref, e, floor))
    assert abs(loss - ref) <= 1e-2 * ref, (loss, ref)
    assert e["bucket_rel_l2"] <= 1.5 * floor["bucket_rel_l2"], (e, floor)
    assert e["tensor_rel_l2_max"] <= 1.5 * floor["tensor_rel_l2_max"], (e, floor)


def test_grad_accumulation_and_generic_autograd():
    fx = load_golden("train_tiny")
    _, _, g1, est = _step(fx, torch.float16)
    # a second backward accumulates into the same flat bucket (autograd semantics)
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    c = lambda k: fx[k].cuda()
    t = 1 - torch.cos(c("t_rand") * 0.5 * 3.14159265359)
    loss, _ = cfm._loss_with_noise(c("x1"), c("mask"), c("mu"), c("spks"), c("cond"), None, t, c("z"), c("cfg_rand") > 0.2)
    (0.5 * loss).backward()
    for k, p in est.named_parameters():
        if p.requires_grad:
            assert torch.allclose(p.grad.cpu(), 1.5 * g1[k], atol=1e-7 + 2e-3 * float(g1[k].abs().max())), k
    # generic path: estimator(...) under autograd with an arbitrary downstream loss
    est.zero_grad(set_to_none=True)
    out = est(c("y") if "y" in fx else c("x1"), c("mask"), c("mu"), t.view(-1), c("spks"), c("cond"))
    assert out.requires_grad
    (out ** 2).mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in est.parameters() if p.requires_grad)
    assert sum(float(p.grad.abs().sum()) for p in est.parameters() if p.requires_grad) > 0


@pytest.mark.parametrize("streams", [2, 3])
def test_sharded_streams_match_single_stream(streams):
    """Concurrent batch shards (one handle + stream per shard, global loss normaliser) == single stream."""
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    fx = load_golden("train_tiny_prompt")          # B = 3, prompt isolation, boundary weights
    ref_loss, _, ref_grads, _ = _step(fx, torch.float16)
    est, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    est = est.cuda().train()
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    cfm.num_streams = streams
    c = lambda k: fx[k].cuda()
    t = 1 - torch.cos(c("t_rand") * 0.5 * 3.14159265359)
    loss, y = cfm._loss_with_noise(c("x1"), c("mask"), c("mu"), c("spks"), c("cond"), fx["prompt_lens"], t, c("z"),
                                   c("cfg_rand") > 0.2)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)
    assert torch.allclose(y.cpu(), fx["y"], atol=1e-5)
    for k, p in est.named_parameters():
        if p.requires_grad:
            g = p.grad.cpu()
            assert torch.allclose(g, ref_grads[k], atol=1e-7 + 1e-3 * float(ref_grads[k].abs().max())), k
    # golden tolerance as well
    assert abs(loss.item() - float(fx["loss"])) <= 1e-2 * float(fx["loss"])


def test_long_ragged_utterances_T1500():
    """BASELINE configs[4] shape class: T = 1500 padded frames with ragged lengths (L = 1500 / 750
    attention, 12 / 6 key blocks), tiny estimator so the CPU oracle finishes in seconds."""
    from oracle import flow_oracle as O
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    from tests.helpers import lora_scaling_of
    B, T = 2, 1500
    est, sd, _ = build_estimator(1, 1, lora_r=8)
    g = torch.Generator().manual_seed(42)
    x1, mu = torch.randn(B, 80, T, generator=g), torch.randn(B, 80, T, generator=g)
    spks, cond = torch.randn(B, 80, generator=g), torch.zeros(B, 80, T)
    lengths = torch.tensor([1500, 611])
    mask = (~O.make_pad_mask(lengths, T)).float().unsqueeze(1)
    t_rand, z, cfg = torch.rand(B, 1, 1, generator=g), torch.randn(B, 80, T, generator=g), torch.tensor([0.9, 0.5])
    P = {k: v.clone().requires_grad_(k.endswith(("lora_A", "lora_B"))) for k, v in sd.items()}
    ref_loss, _, _ = O.cfm_compute_loss(P, x1, mask, mu, spks, cond, None, t_rand, z, cfg, lora_scaling=lora_scaling_of(sd))
    ref_loss.backward()
    est = est.cuda().train()
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    t = 1 - torch.cos(t_rand * 0.5 * 3.14159265359)
    loss, _ = cfm._loss_with_noise(x1.cuda(), mask.cuda(), mu.cuda(), spks.cuda(), cond.cuda(), None, t.cuda(), z.cuda(),
                                   (cfg > 0.2).cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item())
    num = den = 0.0
    for k, p in est.named_parameters():
        if p.requires_grad:
            num += float((p.grad.cpu() - P[k].grad).double().pow(2).sum())
            den += float(P[k].grad.double().pow(2).sum())
    assert (num / den) ** 0.5 <= 1e-2, (num / den) ** 0.5


@pytest.mark.parametrize("name,dtype,tol", [("inputgrads_tiny_prompt", torch.float16, 1e-2),
                                           ("inputgrads_c1", torch.float16, 1e-2),
                                           ("inputgrads_c1", torch.bfloat16, 6e-2)])
def test_input_gradients_vs_reference(name, dtype, tol):
    """dL/dmu, dL/dspks, dL/dcond through cvflow_estimator_backward_inputs vs the reference's autograd
    (golden vectors): rel-L2 <= 1e-2 with fp16 operands (bf16: its own noise floor, as for the LoRA grads);
    exact zeros on padded frames; LoRA gradients unchanged by asking for the input gradients."""
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    ig = load_golden(name)
    fx = load_golden(ig["src"])
    _, _, ref_lora_grads, _ = _step(fx, dtype)
    est, sd, _ = build_estimator(fx["n_blocks"], fx["n_mid"], lora_r=8)
    est = est.cuda().train()
    est.cvflow_dtype = dtype
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
    c = lambda k: fx[k].cuda()
    mu, spks, cond = (c(k).requires_grad_(True) for k in ("mu", "spks", "cond"))
    t = 1 - torch.cos(c("t_rand") * 0.5 * 3.14159265359)
    loss, _ = cfm._loss_with_noise(c("x1"), c("mask"), mu, spks, cond, fx["prompt_lens"], t, c("z"), c("cfg_rand") > 0.2)
    (2.0 * loss).backward()                  # the upstream factor must reach the input gradients too
    assert abs(loss.item() - float(ig["loss"])) <= 1e-2 * float(ig["loss"])
    for got, key in ((mu.grad, "dmu"), (spks.grad, "dspks"), (cond.grad, "dcond")):
        want = 2.0 * ig[key]
        rel = float((got.cpu() - want).double().norm() / want.double().norm())
        assert rel <= tol, (key, rel)
    pad = (fx["mask"].expand_as(fx["mu"]) == 0)
    assert float(mu.grad.cpu()[pad].abs().sum()) == 0.0 and float(cond.grad.cpu()[pad].abs().sum()) == 0.0
    for k, p in est.named_parameters():
        if p.requires_grad:
            assert torch.allclose(p.grad.cpu(), 2.0 * ref_lora_grads[k],
                                  atol=1e-7 + 2e-3 * float(ref_lora_grads[k].abs().max())), k


def test_input_gradients_generic_autograd():
    """ConditionalDecoder.forward under autograd with an arbitrary loss: dL/dx, dL/dmu, dL/dspks, dL/dcond vs the oracle."""
    from oracle import flow_oracle as O
    from tests.helpers import lora_scaling_of
    fx = load_golden("train_tiny")
    est, sd, _ = build_estimator(1, 1, lora_r=8)
    t = 1 - torch.cos(fx["t_rand"] * 0.5 * 3.14159265359)
    leaves = [fx[k].clone().requires_grad_(True) for k in ("y", "mu", "spks", "cond")]
    out = O.estimator_forward(sd, leaves[0], fx["mask"], leaves[1], t.view(-1), leaves[2], leaves[3],
                              lora_scaling=lora_scaling_of(sd))
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * gout).sum().backward()
    est = est.cuda().train()
    dl = [fx[k].cuda().requires_grad_(True) for k in ("y", "mu", "spks", "cond")]
    o = est(dl[0], fx["mask"].cuda(), dl[1], t.view(-1).cuda(), dl[2], dl[3])
    (o * gout.cuda()).sum().backward()
    assert torch.allclose(o.detach().cpu(), out.detach(), atol=2e-2, rtol=2e-2)
    for a, b, key in zip(dl, leaves, ("dx", "dmu", "dspks", "dcond")):
        rel = float((a.grad.cpu() - b.grad).double().norm() / b.grad.double().norm())
        assert rel <= 1e-2, (key, rel)
