"""End-to-end through the reference-shaped Python API on the GPU: build_flow_model -> LoRA inject ->
JointLLMFlowModel('flow_only') training steps with the DP trainer -> merged checkpoint -> strict load
into a fresh stock-layout model -> Euler inference; plus the train_joint CLI on synthetic batches."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=1234):
    from cosyvoice_lora_finetune_framework_b200 import flow_model, lora, utils
    from cosyvoice_lora_finetune_framework_b200.llm_flow_model import JointLLMFlowModel
    utils.set_all_random_seed(seed)
    flow = flow_model.build_flow_model(None, 'cpu', decoder_n_blocks=1, decoder_num_mid_blocks=2)
    stats = lora.apply_lora_to_model(flow, r=8, lora_alpha=16, lora_dropout=0.0, target_modules=['to_q', 'to_k', 'to_v'])
    assert stats['replaced_layers'] == 18          # (2 + 2 + 2) stages x 1 block x q,k,v
    return JointLLMFlowModel(None, flow.cuda(), 'flow_only', 1.0, 1.0, True)


def test_flow_only_training_reduces_loss_and_merges(tmp_path):
    from cosyvoice_lora_finetune_framework_b200 import flow_model, lora
    from cosyvoice_lora_finetune_framework_b200.train_joint import synthetic_batches
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
    model = _model()
    model.eval()        # deterministic encoder (no dropout); the estimator itself has none
    model.flow.decoder.estimator.train()
    batch = next(synthetic_batches(1, 4, max_feat_len=96, seed=3))
    trainer = FlowLoRATrainer(model.flow.decoder, lr=2e-3, max_grad_norm=1.0)
    dev = torch.device('cuda')
    losses = []
    for step in range(12):
        torch.manual_seed(0)                        # same (t, z) every step: a fixed objective to descend
        out = model(batch, dev)
        assert set(out) == {'flow_loss', 'loss'}
        out['loss'].backward()
        trainer.micro = 1
        trainer.optimizer_step()
        losses.append(float(out['loss']))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0] * 0.97, losses
    assert int(trainer.found_inf.item()) == 0
    # merged checkpoint: stock key layout, loads strictly, reproduces the LoRA model's inference
    flow = model.flow
    flow.eval()
    tok = batch['speech_token'][:1, :40].cuda()
    emb = batch['embedding'][:1].cuda()
    args = (tok, torch.tensor([40]).cuda(), torch.zeros(1, 0, dtype=torch.long).cuda(), torch.tensor([0]).cuda(),
            torch.zeros(1, 0, 80).cuda(), torch.tensor([0]).cuda(), emb)
    torch.manual_seed(5)
    mel_lora, _ = flow.inference(*args)
    merged = lora.get_merged_state_dict(flow)
    path = str(tmp_path / 'flow_merged_flow_only.pt')
    torch.save(merged, path)
    fresh = flow_model.build_flow_model(None, 'cpu', decoder_n_blocks=1, decoder_num_mid_blocks=2)
    assert list(fresh.state_dict().keys()) != [] and set(fresh.state_dict().keys()) == set(merged.keys())
    fresh.load_state_dict(torch.load(path), strict=True)
    fresh = fresh.cuda().eval()
    torch.manual_seed(5)
    mel_merged, _ = fresh.inference(*args)
    assert mel_lora.shape == mel_merged.shape == (1, 80, int(40 / 50 * 22050 / 256))
    assert (mel_lora - mel_merged).abs().max().item() <= 2e-2 * mel_lora.abs().max().item()


def test_requires_grad_on_prepared_tensors():
    """mu / spks / cond that require grad get their gradients (modules upstream of the estimator can be trained);
    a target mel that requires grad is refused loudly."""
    model = _model()
    est = model.flow.decoder
    x = torch.randn(1, 80, 16, device='cuda')
    mu = torch.randn(1, 80, 16, device='cuda', requires_grad=True)
    spk = torch.randn(1, 80, device='cuda', requires_grad=True)
    est.training_cfg_rate = 0.0
    loss, _ = est.compute_loss(x, torch.ones(1, 1, 16, device='cuda'), mu, spk, cond=torch.zeros(1, 80, 16, device='cuda'))
    loss.backward()
    for g in (mu.grad, spk.grad):
        assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    with pytest.raises(NotImplementedError):
        est.compute_loss(x.clone().requires_grad_(True), torch.ones(1, 1, 16, device='cuda'), mu.detach(), spk.detach(),
                         cond=torch.zeros(1, 80, 16, device='cuda'))


def test_train_joint_cli_synthetic(tmp_path):
    from cosyvoice_lora_finetune_framework_b200 import config, train_joint
    old = dict(config.JOINT_TRAINING_CONFIG)
    config.JOINT_TRAINING_CONFIG.update(accumulate_grad_batches=2, max_feat_len=64)
    try:
        train_joint.main(['--mode', 'flow_only', '--epochs', '1', '--batch-size', '2', '--synthetic', '4',
                          '--output-dir', str(tmp_path)])
    finally:
        config.JOINT_TRAINING_CONFIG.clear()
        config.JOINT_TRAINING_CONFIG.update(old)
    ck = torch.load(os.path.join(str(tmp_path), 'joint_flow_only_last.ckpt'), map_location='cpu', weights_only=False)
    assert any(k.startswith('model.flow.decoder.estimator.') and k.endswith('lora_A') for k in ck['state_dict'])
    assert any(k.startswith('model.flow.encoder.') and k.endswith('lora_A') for k in ck['state_dict'])
    merged = torch.load(os.path.join(str(tmp_path), 'flow_merged_flow_only.pt'), map_location='cpu')
    assert not any('lora_' in k or 'original_layer' in k for k in merged)


def test_graphed_step_matches_eager_step():
    """The whole-step CUDA graph must produce the same parameters as eager launches (same seeds)."""
    from cosyvoice_lora_finetune_framework_b200.train_joint import synthetic_batches
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
    results = []
    for graphed in (False, True):
        model = _model(seed=7)
        model.eval()
        cfm = model.flow.decoder
        cfm.estimator.train()
        tr = FlowLoRATrainer(cfm, lr=1e-3, warmup_steps=2, total_steps=10)
        g = torch.Generator().manual_seed(1)
        B, T = 3, 72
        x1, mu = torch.randn(B, 80, T, generator=g).cuda(), torch.randn(B, 80, T, generator=g).cuda()
        spks, cond = torch.randn(B, 80, generator=g).cuda(), torch.zeros(B, 80, T).cuda()
        mask = torch.ones(B, 1, T).cuda()
        mask[1, :, 50:] = 0
        torch.manual_seed(123)
        torch.cuda.manual_seed_all(123)
        losses = []
        for i in range(5):
            if graphed:
                losses.append(float(tr.train_step_graphed(x1, mask, mu, spks, cond)))
            else:
                losses.append(float(tr.train_step(x1, mask, mu, spks, cond)))
        results.append((losses, tr.ne.param_bucket.clone(), tr.step_count))
    (l0, p0, s0), (l1, p1, s1) = results
    assert all(torch.isfinite(torch.tensor(l0 + l1)))
    # one call == one optimiser step in both forms: the warm-up executions that precede the capture are rolled back
    # (parameters, Adam moments, device-side step counter, RNG), so the two trajectories coincide
    assert s1 == s0 == 5
    assert torch.allclose(torch.tensor(l1), torch.tensor(l0), rtol=2e-3, atol=1e-5), (l0, l1)
    assert torch.allclose(p1, p0, rtol=1e-3, atol=2e-5), float((p1 - p0).abs().max())


def test_graphed_accumulation_matches_eager_accumulation():
    """accumulate_grad_batches > 1 (the reference trains with batch 1 x accumulate 16, config.py): the micro-step graph
    replayed per micro-batch plus the optimiser-tail graph after every window == eager micro_step / optimizer_step."""
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
    results = []
    g = torch.Generator().manual_seed(2)
    B, T = 2, 72
    batches = [(torch.randn(B, 80, T, generator=g).cuda(), torch.randn(B, 80, T, generator=g).cuda(),
                torch.randn(B, 80, generator=g).cuda()) for _ in range(3)]
    cond, mask = torch.zeros(B, 80, T).cuda(), torch.ones(B, 1, T).cuda()
    mask[1, :, 40:] = 0
    for graphed in (False, True):
        model = _model(seed=9)
        model.eval()
        cfm = model.flow.decoder
        cfm.estimator.train()
        tr = FlowLoRATrainer(cfm, lr=1e-3, accumulate=3)
        torch.manual_seed(321)
        torch.cuda.manual_seed_all(321)
        losses = []
        for w in range(2):                       # two optimiser steps of three micro-batches each
            for x1, mu, spks in batches:
                fn = tr.train_step_graphed if graphed else tr.train_step
                losses.append(float(fn(x1, mask, mu, spks, cond)))
        results.append((losses, tr.ne.param_bucket.clone(), tr.step_count, int(tr.opt_state[0].item())))
    (l0, p0, s0, d0), (l1, p1, s1, d1) = results
    assert s0 == s1 == 2 and d0 == d1 == 2
    assert torch.allclose(torch.tensor(l1), torch.tensor(l0), rtol=2e-3, atol=1e-5), (l0, l1)
    assert torch.allclose(p1, p0, rtol=1e-3, atol=2e-5), float((p1 - p0).abs().max())


def test_optimizer_schedule_state_and_resume():
    """Device-side step counter / LR schedule (LambdaLR semantics: the first warm-up step runs with lr = 0), skipped
    steps on a non-finite gradient norm, and trainer.state_dict() / load_state_dict() round trip."""
    from cosyvoice_lora_finetune_framework_b200.train_joint import synthetic_batches  # noqa: F401
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer, lr_lambda
    model = _model(seed=11)
    model.eval()
    cfm = model.flow.decoder
    cfm.estimator.train()
    tr = FlowLoRATrainer(cfm, lr=1e-3, warmup_steps=4, total_steps=20)
    g = torch.Generator().manual_seed(2)
    B, T = 2, 48
    x1, mu = torch.randn(B, 80, T, generator=g).cuda(), torch.randn(B, 80, T, generator=g).cuda()
    spks, cond, mask = torch.randn(B, 80, generator=g).cuda(), torch.zeros(B, 80, T).cuda(), torch.ones(B, 1, T).cuda()
    p_init = tr.ne.param_bucket.clone()
    tr.train_step(x1, mask, mu, spks, cond)
    torch.cuda.synchronize()
    # step 1 of a warm-up: lambda(0) = 0 -> parameters unchanged (AdamW decay is lr-scaled too), counter advanced
    assert torch.equal(tr.ne.param_bucket, p_init)
    assert tr.opt_state.tolist() == [1, 0] and float(tr.hyper[0]) == 0.0
    tr.train_step(x1, mask, mu, spks, cond)
    assert abs(float(tr.hyper[0]) - 1e-3 * lr_lambda(1, 4, 20, 1e-3, 1e-6)) < 1e-9
    assert not torch.equal(tr.ne.param_bucket, p_init)
    # a non-finite gradient: the step is skipped, neither the counter nor the parameters move, the skip is counted
    p_before = tr.ne.param_bucket.clone()
    loss, _ = cfm.compute_loss(x1, mask, mu, spks, cond=cond)
    loss.backward()
    tr.ne.grad_bucket[0] = float("inf")
    tr.micro = 1
    tr.optimizer_step()
    torch.cuda.synchronize()
    assert tr.opt_state.tolist() == [2, 1] and int(tr.found_inf.item()) == 1
    assert torch.equal(tr.ne.param_bucket, p_before)
    scale0 = tr.ne.loss_scale
    assert tr.poll_overflow() == 1 and tr.step_count == 2 and int(tr.found_inf.item()) == 0
    assert tr.ne.loss_scale == max(1.0, scale0 * 0.5)
    # resume: a fresh trainer with the saved optimiser state continues exactly like the original
    sd_model = {k: v.clone() for k, v in model.state_dict().items()}
    sd_opt = tr.state_dict()
    torch.manual_seed(5)
    tr.train_step(x1, mask, mu, spks, cond)
    want = tr.ne.param_bucket.clone()
    model2 = _model(seed=12)
    model2.eval()
    model2.load_state_dict(sd_model)
    model2.flow.decoder.estimator.train()
    tr2 = FlowLoRATrainer(model2.flow.decoder, lr=1e-3, warmup_steps=4, total_steps=20)
    tr2.load_state_dict(sd_opt)
    assert tr2.step_count == 2
    torch.manual_seed(5)
    tr2.train_step(x1, mask, mu, spks, cond)
    assert torch.allclose(tr2.ne.param_bucket, want, rtol=1e-4, atol=1e-7), float((tr2.ne.param_bucket - want).abs().max())
