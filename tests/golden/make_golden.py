"""Generate golden vectors by running the REAL reference implementation (read-only checkout at
/root/reference) on seeded synthetic weights and inputs. Run in the build container only:

    python tests/golden/make_golden.py               # train_*, estimator_*, euler_* (the original set)
    python tests/golden/make_golden.py inputgrads    # inputgrads_*: dL/dmu, dL/dspks, dL/dcond of compute_loss
    python tests/golden/make_golden.py flowmodel     # flowmodel_tiny: MaskedDiffWithXvec.forward + backward, encoder LoRA
    python tests/golden/make_golden.py dropout       # dropout_tiny_prompt: lora_dropout > 0 with preset keep masks
    python tests/golden/make_golden.py bench         # the BENCHMARKED shapes: train_c3 (300M, 32 x 400 ragged, + the reference's
                                                     # own bf16 / fp16 autocast error at that shape), euler_c2 (300M, T=700,
                                                     # P=200, 10 steps), train_tiny_padprompt (boundary window running into the
                                                     # padding), dropout_c1 (300M, lora_dropout = 0.05 with preset masks)

    python tests/golden/make_golden.py flowinfer     # flowmodel_infer_tiny: MaskedDiffWithXvec.inference / inference_like_training
    python tests/golden/make_golden.py regulator     # regulator_tiny / regulator_c3: InterpolateRegulator forward, inference and
                                                     # input gradient, real reference modules.py:800-837

The GPU box has no /root/reference; tests read the committed .pt files. Weights are not stored:
they are a pure function of (parameter name, shape, seed) - oracle.flow_oracle.synth_tensor - and
each fixture records a checksum so a drift in that generator is detected.
"""
import os
import sys

import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CVFLOW_REF", "/root/reference/cosyvoice_flow_finetune")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import flow_model as ref_flow          # noqa: E402  (reference)
import lora as ref_lora                # noqa: E402
import modules as ref_modules          # noqa: E402
import utils as ref_utils              # noqa: E402

from oracle import flow_oracle as O    # noqa: E402
from cosyvoice_lora_finetune_framework_b200 import modules as our_modules  # noqa: E402

torch.set_num_threads(os.cpu_count())
WSEED = 1234


def build_ref(n_blocks, n_mid, r=8, alpha=16, targets=('to_q', 'to_k', 'to_v', 'to_out')):
    est = ref_modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0,
                                         attention_head_dim=64, n_blocks=n_blocks, num_mid_blocks=n_mid,
                                         num_heads=8, act_fn='gelu')
    stats = None
    if r:
        stats = ref_lora.apply_lora_to_model(est, r=r, lora_alpha=alpha, lora_dropout=0.0,
                                             target_modules=list(targets))
    spec = {k: tuple(v.shape) for k, v in est.state_dict().items()}
    sd = O.synth_state_dict(spec, WSEED)
    est.load_state_dict(sd, strict=True)
    cfm = ref_flow.ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, sigma_min=1e-6, t_scheduler='cosine',
                                  training_cfg_rate=0.2, inference_cfg_rate=0.7, estimator=est)
    return cfm, sd, stats


def wsum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def train_case(name, n_blocks, n_mid, B, T, lengths, prompt_lens, data_seed, step_seed, keep_grads):
    cfm, sd, stats = build_ref(n_blocks, n_mid)
    cfm.train()
    g = torch.Generator().manual_seed(data_seed)
    x1 = torch.randn(B, 80, T, generator=g)
    mu = torch.randn(B, 80, T, generator=g)
    spks = torch.randn(B, 80, generator=g)
    cond = torch.zeros(B, 80, T)
    if prompt_lens:
        for i, p in enumerate(prompt_lens):
            cond[i, :, :p] = x1[i, :, :p]
    lens = torch.tensor(lengths)
    mask = (~ref_utils.make_pad_mask(lens, T)).float().unsqueeze(1)
    # replicate the reference's three draws (flow_model.py:146,151,159) for the fixture
    torch.manual_seed(step_seed)
    t_rand = torch.rand([B, 1, 1])
    z = torch.randn_like(x1)
    cfg_rand = torch.rand(B)
    torch.manual_seed(step_seed)
    loss, y = cfm.compute_loss(x1, mask, mu, spks, cond=cond, prompt_lens=prompt_lens)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in cfm.estimator.named_parameters() if p.grad is not None}
    # pred via a second forward with identical draws
    with torch.no_grad():
        t = 1 - torch.cos(t_rand * 0.5 * 3.14159265359)
        keep = (cfg_rand > 0.2)
        cfm.estimator.prompt_isolation_len = max(prompt_lens) if prompt_lens else 0
        pred = cfm.estimator(y, mask, mu * keep.view(-1, 1, 1), t.view(B), spks * keep.view(-1, 1),
                             cond * keep.view(-1, 1, 1))
        cfm.estimator.prompt_isolation_len = 0
    fx = dict(kind="train", n_blocks=n_blocks, n_mid=n_mid, wseed=WSEED, wsum=wsum(sd), lora_stats=stats,
              x1=x1, mu=mu, spks=spks, cond=cond, mask=mask, lengths=lens, prompt_lens=prompt_lens,
              t_rand=t_rand, z=z, cfg_rand=cfg_rand, loss=loss.detach(), y=y.detach(), pred=pred,
              grad_norms={k: float(v.norm()) for k, v in grads.items()},
              grad_total_norm=float(torch.sqrt(sum(v.double().pow(2).sum() for v in grads.values()))),
              grads={k: v for k, v in grads.items() if keep_grads(k)})
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "loss", float(loss), "gradnorm", fx["grad_total_norm"], "ngrads", len(grads), "kept", len(fx["grads"]))


def input_grad_case(name, src):
    """dL/dmu, dL/dspks, dL/dcond of the reference's compute_loss (same inputs and draws as the fixture `src`):
    what the encoder / speaker-affine LoRA layers upstream of the estimator receive (config.py:207-216)."""
    fx = torch.load(os.path.join(HERE, src + ".pt"), map_location="cpu", weights_only=False)
    cfm, sd, _ = build_ref(fx["n_blocks"], fx["n_mid"])
    assert abs(wsum(sd) - fx["wsum"]) < 1e-6 * fx["wsum"]
    cfm.train()
    mu = fx["mu"].clone().requires_grad_(True)
    spks = fx["spks"].clone().requires_grad_(True)
    cond = fx["cond"].clone().requires_grad_(True)
    step_seed = {"train_tiny": 7, "train_tiny_prompt": 8, "train_c1": 7, "train_c1_prompt": 7}[src]
    torch.manual_seed(step_seed)
    loss, _ = cfm.compute_loss(fx["x1"], fx["mask"], mu, spks, cond=cond, prompt_lens=fx["prompt_lens"])
    assert float(loss) == float(fx["loss"]), (float(loss), float(fx["loss"]))
    loss.backward()
    torch.save(dict(kind="input_grads", src=src, loss=loss.detach(), dmu=mu.grad.clone(), dspks=spks.grad.clone(),
                    dcond=cond.grad.clone()), os.path.join(HERE, name + ".pt"))
    print(name, "loss", float(loss), "|dmu|", float(mu.grad.norm()), "|dspks|", float(spks.grad.norm()), "|dcond|",
          float(cond.grad.norm()))


def dropout_masks(n_tbs, rows, p, seed):
    """Explicit keep masks [n_tbs][3][rows][256] (uint8) shared by the reference run, the oracle and the CUDA path."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n_tbs, 3, rows, 256, generator=g) >= p).to(torch.uint8)


def attention_blocks(est, prefix=""):
    """(state-dict prefix, module) of every BasicTransformerBlock in execution order (= the CUDA path's block index)."""
    out = []
    groups = ([("down_blocks.%d" % i, b) for i, b in enumerate(est.down_blocks)] +
              [("mid_blocks.%d" % i, b) for i, b in enumerate(est.mid_blocks)] +
              [("up_blocks.%d" % i, b) for i, b in enumerate(est.up_blocks)])
    for name, stage in groups:
        for j, tb in enumerate(stage[1]):
            out.append(("%s%s.1.%d" % (prefix, name, j), tb))
    return out


class FixedDropout(torch.nn.Module):
    """Stands in for a LoRALinear's nn.Dropout: multiplies by a preset keep mask / (1 - p)."""

    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        b, l, c = x.shape
        return x * (self.keep[: b * l].view(b, l, c).to(x.dtype) / (1.0 - self.p))


def dropout_case(name, src, p, mask_seed):
    """compute_loss + backward of the REAL reference with lora_dropout = p on the estimator's q/k/v LoRA layers, the
    dropout draws replaced by preset masks (one independent mask per LoRALinear, like nn.Dropout)."""
    fx = torch.load(os.path.join(HERE, src + ".pt"), map_location="cpu", weights_only=False)
    cfm, sd, _ = build_ref(fx["n_blocks"], fx["n_mid"], targets=('to_q', 'to_k', 'to_v'))
    cfm.train()
    B, _, T = fx["x1"].shape
    blocks = attention_blocks(cfm.estimator)
    keep = dropout_masks(len(blocks), B * T, p, mask_seed)
    for i, (_, tb) in enumerate(blocks):
        for j, pn in enumerate(("to_q", "to_k", "to_v")):
            getattr(tb.attn1, pn).lora_dropout = FixedDropout(keep[i, j], p)
    step_seed = {"train_tiny": 7, "train_tiny_prompt": 8, "train_c1": 7, "train_c1_prompt": 7}[src]
    torch.manual_seed(step_seed)
    loss, _ = cfm.compute_loss(fx["x1"], fx["mask"], fx["mu"], fx["spks"], cond=fx["cond"], prompt_lens=fx["prompt_lens"])
    loss.backward()
    grads = {k: q.grad.clone() for k, q in cfm.estimator.named_parameters() if q.grad is not None}
    torch.save(dict(kind="lora_dropout", src=src, p=p, mask_seed=mask_seed, n_tbs=len(blocks), rows=B * T,
                    keep_sum=int(keep.sum()), loss=loss.detach(), grads=grads), os.path.join(HERE, name + ".pt"))
    print(name, "loss", float(loss), "(no dropout: %g)" % float(fx["loss"]), "ngrads", len(grads), "keep rate",
          float(keep.float().mean()))


# ------------------------------------------------------------------------------------------------------
# Fixtures at the benchmarked shapes. The big input tensors are NOT stored: they are regenerated from the
# seeds with the same torch.Generator call sequence (bench.py::make_batch), and the fixture keeps checksums.
# ------------------------------------------------------------------------------------------------------
def bench_batch(B, T, seed):
    """Exactly bench.py::make_batch (configs[2] synthetic batch)."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(B, 80, T, generator=g)
    mu = torch.randn(B, 80, T, generator=g)
    spks = torch.randn(B, 80, generator=g)
    cond = torch.zeros(B, 80, T)
    lens = torch.randint(int(0.6 * T) + 1, T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(1)
    return x1, mu, spks, cond, mask, lens


def step_draws(B, T, step_seed):
    torch.manual_seed(step_seed)
    t_rand = torch.rand([B, 1, 1])
    z = torch.randn(B, 80, T)
    cfg_rand = torch.rand(B)
    return t_rand, z, cfg_rand


def csum(t):
    return [float(t.double().sum()), float(t.double().abs().sum())]


def _grad_errors(grads, ref):
    num = sum((grads[k] - ref[k]).double().pow(2).sum() for k in ref)
    den = sum(ref[k].double().pow(2).sum() for k in ref)
    per = sorted(float((grads[k] - ref[k]).double().norm() / (ref[k].double().norm() + 1e-30)) for k in ref)
    return dict(bucket_rel_l2=float((num / den).sqrt()), tensor_rel_l2_median=per[len(per) // 2], tensor_rel_l2_max=per[-1])


def train_bench_case(name, B, T, batch_seed, step_seed, keep_grads, autocast=True):
    cfm, sd, stats = build_ref(4, 12)
    cfm.train()
    x1, mu, spks, cond, mask, lens = bench_batch(B, T, batch_seed)
    t_rand, z, cfg_rand = step_draws(B, T, step_seed)
    torch.manual_seed(step_seed)
    loss, y = cfm.compute_loss(x1, mask, mu, spks, cond=cond, prompt_lens=None)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in cfm.estimator.named_parameters() if p.grad is not None}
    with torch.no_grad():
        t = 1 - torch.cos(t_rand * 0.5 * 3.14159265359)
        assert torch.equal(y, (1 - (1 - 1e-6) * t) * z + t * x1)      # the recorded draws are the ones compute_loss made
    floors = {}
    if autocast:      # the reference's own reduced-precision error at this exact shape (its noise floor)
        for dn, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
            for p in cfm.estimator.parameters():
                p.grad = None
            torch.manual_seed(step_seed)
            with torch.autocast("cpu", dtype=dt):
                l2, _ = cfm.compute_loss(x1, mask, mu, spks, cond=cond, prompt_lens=None)
            scale = 4096.0 if dn == "fp16" else 1.0     # fp16 gradients need the GradScaler's loss scale ('16-mixed')
            (l2 * scale).backward()
            g2 = {k: p.grad.float().clone() / scale for k, p in cfm.estimator.named_parameters() if p.grad is not None}
            floors[dn] = dict(loss_rel=abs(float(l2) - float(loss)) / abs(float(loss)), **_grad_errors(g2, grads))
            print(name, "reference under", dn, "autocast:", floors[dn])
    fx = dict(kind="train_bench", n_blocks=4, n_mid=12, wseed=WSEED, wsum=wsum(sd), lora_stats=stats, B=B, T=T,
              batch_seed=batch_seed, step_seed=step_seed, lengths=lens,
              checks=dict(x1=csum(x1), mu=csum(mu), spks=csum(spks), z=csum(z), t_rand=csum(t_rand), cfg_rand=csum(cfg_rand),
                          y=csum(y.detach())),
              loss=loss.detach(), grad_norms={k: float(v.norm()) for k, v in grads.items()},
              grad_total_norm=float(torch.sqrt(sum(v.double().pow(2).sum() for v in grads.values()))),
              grads={k: v for k, v in grads.items() if keep_grads(k)}, ref_autocast=floors)
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "loss", float(loss), "gradnorm", fx["grad_total_norm"], "ngrads", len(grads), "kept", len(fx["grads"]),
          "valid frames", int(lens.sum()))


def euler_bench_case(name, T, prompt, n_steps, seed):
    """bench.py's inference leg (configs[1]): same generator sequence for mu / spks / cond prompt."""
    cfm, sd, _ = build_ref(4, 12, r=0)
    cfm.eval()
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(1, 80, T, generator=g)
    spks = torch.randn(1, 80, generator=g)
    cond = torch.zeros(1, 80, T)
    cond[:, :, :prompt] = torch.randn(1, 80, prompt, generator=g)
    mask = torch.ones(1, 1, T)
    torch.manual_seed(seed + 1)
    z = torch.randn_like(mu)
    torch.manual_seed(seed + 1)
    mel, cache = cfm(mu=mu.clone(), mask=mask, n_timesteps=n_steps, temperature=1.0, spks=spks, cond=cond,
                     prompt_len=prompt, cache=None)
    floors = {}
    for dn, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        torch.manual_seed(seed + 1)
        with torch.autocast("cpu", dtype=dt):
            m2, _ = cfm(mu=mu.clone(), mask=mask, n_timesteps=n_steps, temperature=1.0, spks=spks, cond=cond,
                        prompt_len=prompt, cache=None)
        floors[dn] = dict(max_abs=float((m2.float() - mel).abs().max()), mean_abs=float((m2.float() - mel).abs().mean()))
        print(name, "reference under", dn, "autocast:", floors[dn], "abs-max of mel", float(mel.abs().max()))
    torch.save(dict(kind="euler_bench", n_blocks=4, n_mid=12, wseed=WSEED, wsum=wsum(sd), T=T, prompt=prompt, n_steps=n_steps,
                    seed=seed, checks=dict(mu=csum(mu), spks=csum(spks), cond=csum(cond), z=csum(z)), mel=mel.clone(),
                    cache=cache.clone(), ref_autocast=floors), os.path.join(HERE, name + ".pt"))
    print(name, "mel sum", float(mel.sum()), "absmax", float(mel.abs().max()), tuple(cache.shape))


def dropout_bench_case(name, src, p, mask_seed, keep_grads):
    """300M-scale lora_dropout fixture (the reference's default p = 0.05): like dropout_case, full grads only for a few
    layers, masks regenerated from the seed by the test."""
    fx = torch.load(os.path.join(HERE, src + ".pt"), map_location="cpu", weights_only=False)
    cfm, sd, _ = build_ref(fx["n_blocks"], fx["n_mid"], targets=('to_q', 'to_k', 'to_v'))
    cfm.train()
    B, _, T = fx["x1"].shape
    blocks = attention_blocks(cfm.estimator)
    keep = dropout_masks(len(blocks), B * T, p, mask_seed)
    for i, (_, tb) in enumerate(blocks):
        for j, pn in enumerate(("to_q", "to_k", "to_v")):
            getattr(tb.attn1, pn).lora_dropout = FixedDropout(keep[i, j], p)
    torch.manual_seed(7)
    loss, _ = cfm.compute_loss(fx["x1"], fx["mask"], fx["mu"], fx["spks"], cond=fx["cond"], prompt_lens=fx["prompt_lens"])
    loss.backward()
    grads = {k: q.grad.clone() for k, q in cfm.estimator.named_parameters() if q.grad is not None}
    torch.save(dict(kind="lora_dropout", src=src, p=p, mask_seed=mask_seed, n_tbs=len(blocks), rows=B * T,
                    keep_sum=int(keep.sum()), loss=loss.detach(), grad_norms={k: float(v.norm()) for k, v in grads.items()},
                    grads={k: v for k, v in grads.items() if keep_grads(k)}), os.path.join(HERE, name + ".pt"))
    print(name, "loss", float(loss), "(no dropout: %g)" % float(fx["loss"]), "ngrads", len(grads), "keep rate",
          float(keep.float().mean()))


FLOW_TINY = dict(encoder_num_blocks=2, decoder_n_blocks=1, decoder_num_mid_blocks=1)
FLOW_TARGETS = ['to_q', 'to_k', 'to_v', 'linear_q', 'linear_k', 'linear_v', 'w_1', 'w_2']     # config.py:207-216


def flow_model_case(name, init_seed, py_seed, step_seed):
    """MaskedDiffWithXvec.forward(batch) + backward of the REAL reference with LoRA on the estimator's attn1 q/k/v
    AND on the Conformer encoder (the reference's flow_lora target list): loss and every LoRA gradient. Weights are
    the seeded random init (bit-identical between the reference and our module tree, asserted here). The encoder runs
    in eval() so that its dropout does not consume the torch RNG; the three draws of compute_loss are recorded."""
    import random
    from cosyvoice_lora_finetune_framework_b200 import flow_model as our_flow, lora as our_lora
    ref_utils.set_all_random_seed(init_seed)
    m = ref_flow.build_flow_model(None, 'cpu', **FLOW_TINY)
    stats = ref_lora.apply_lora_to_model(m, r=8, lora_alpha=16, lora_dropout=0.0, target_modules=FLOW_TARGETS)
    ref_utils.set_all_random_seed(init_seed)
    o = our_flow.build_flow_model(None, 'cpu', **FLOW_TINY)
    our_lora.apply_lora_to_model(o, r=8, lora_alpha=16, lora_dropout=0.0, target_modules=FLOW_TARGETS)
    sa, sb = m.state_dict(), o.state_dict()
    assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa), "seeded init differs"
    m.train()
    m.encoder.eval()
    g = torch.Generator().manual_seed(0)
    B = 3
    batch = dict(speech_token=torch.randint(0, 4096, (B, 40), generator=g), speech_token_len=torch.tensor([40, 33, 21]),
                 speech_feat=torch.randn(B, 70, 80, generator=g) * 2 - 6, speech_feat_len=torch.tensor([70, 57, 36]),
                 embedding=torch.randn(B, 192, generator=g),
                 cross_sample_mel=torch.randn(B, 30, 80, generator=g) * 2 - 6,
                 cross_sample_mel_len=torch.tensor([30, 0, 12]))
    rec = {}
    orig = m.decoder.compute_loss

    def recording(x1, mask, mu, spks, cond=None, prompt_lens=None):
        st = torch.get_rng_state()
        b = mu.shape[0]
        rec.update(t_rand=torch.rand([b, 1, 1]), z=torch.randn_like(x1), cfg_rand=torch.rand(b),
                   prompt_lens=list(prompt_lens) if prompt_lens is not None else None, mu=mu.detach().clone(),
                   spks=spks.detach().clone(), cond=cond.detach().clone(), x1=x1.detach().clone())
        torch.set_rng_state(st)
        return orig(x1, mask, mu, spks, cond=cond, prompt_lens=prompt_lens)

    m.decoder.compute_loss = recording
    random.seed(py_seed)
    torch.manual_seed(step_seed)
    out = m(batch, torch.device('cpu'))
    out['loss'].backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.requires_grad}
    assert all(v is not None for v in grads.values())
    fx = dict(kind="flow_model", arch=FLOW_TINY, targets=FLOW_TARGETS, init_seed=init_seed, py_seed=py_seed, batch=batch,
              wsum=wsum(sa), lora_stats=stats, loss=out['loss'].detach(), grads=grads, **rec)
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    enc = torch.sqrt(sum(v.double().pow(2).sum() for k, v in grads.items() if k.startswith("encoder.")))
    est = torch.sqrt(sum(v.double().pow(2).sum() for k, v in grads.items() if k.startswith("decoder.")))
    print(name, "loss", float(out['loss']), "prompt_lens", rec["prompt_lens"], "ngrads", len(grads), "|enc|", float(enc),
          "|est|", float(est), stats["replaced_layers"])


def flow_model_infer_case(name, init_seed, seed):
    """MaskedDiffWithXvec.inference of the REAL reference (flow_model.py:475-551): prompt + target tokens -> encoder ->
    InterpolateRegulator.inference (prompt | head | middle | tail stretched separately) -> prompt conditioning -> CFM Euler
    solve (step count by length) -> mel without the prompt frames; plus inference_like_training (:553-638). The noise is the
    first draw after torch.manual_seed(seed), so the test can regenerate it."""
    ref_utils.set_all_random_seed(init_seed)
    m = ref_flow.build_flow_model(None, 'cpu', **FLOW_TINY).eval()
    g = torch.Generator().manual_seed(seed)
    n_prompt, n_target = 30, 75
    tok = torch.randint(0, 4096, (1, n_target), generator=g)
    ptok = torch.randint(0, 4096, (1, n_prompt), generator=g)
    pfeat = torch.randn(1, 52, 80, generator=g)
    emb = torch.randn(1, 192, generator=g)
    with torch.no_grad():
        torch.manual_seed(seed)
        mel, cache = m.inference(tok, torch.tensor([n_target]), ptok, torch.tensor([n_prompt]), pfeat, torch.tensor([52]), emb)
        torch.manual_seed(seed + 1)
        mel2 = m.inference_like_training(tok[:, :40], torch.tensor([40]), torch.tensor([68]), emb, prompt_feat=pfeat,
                                         prompt_len=12, n_timesteps=10)
    fx = dict(kind="flow_model_infer", arch=FLOW_TINY, init_seed=init_seed, seed=seed, token=tok, prompt_token=ptok,
              prompt_feat=pfeat, embedding=emb, mel=mel, cache=cache, mel_like_training=mel2, wsum=wsum(m.state_dict()))
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "mel", tuple(mel.shape), "cache", tuple(cache.shape), "like-training", tuple(mel2.shape), float(mel.abs().mean()))


def estimator_case(name, n_blocks, n_mid, Ts, seed):
    """export_onnx.py:34-41,95-116 protocol: batch 2, torch.rand inputs, random T."""
    cfm, sd, _ = build_ref(n_blocks, n_mid, r=0)
    est = cfm.estimator.eval()
    cases = []
    g = torch.Generator().manual_seed(seed)
    for T in Ts:
        x = torch.rand(2, 80, T, generator=g)
        mu = torch.rand(2, 80, T, generator=g)
        t = torch.rand(2, generator=g)
        spks = torch.rand(2, 80, generator=g)
        cond = torch.rand(2, 80, T, generator=g)
        mask = torch.ones(2, 1, T)
        with torch.no_grad():
            out = est(x, mask, mu, t, spks, cond)
        cases.append(dict(T=T, x=x, mu=mu, t=t, spks=spks, cond=cond, mask=mask, out=out))
    torch.save(dict(kind="estimator", n_blocks=n_blocks, n_mid=n_mid, wseed=WSEED, wsum=wsum(sd), cases=cases),
               os.path.join(HERE, name + ".pt"))
    print(name, [float(c["out"].abs().max()) for c in cases])


def euler_case(name, n_blocks, n_mid, T, prompt, n_steps, seed):
    cfm, sd, _ = build_ref(n_blocks, n_mid, r=0)
    cfm.eval()
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(1, 80, T, generator=g)
    spks = torch.randn(1, 80, generator=g)
    cond = torch.zeros(1, 80, T)
    cond[:, :, :prompt] = torch.randn(1, 80, prompt, generator=g)
    mask = torch.ones(1, 1, T)
    torch.manual_seed(seed + 1)
    z = torch.randn_like(mu)
    torch.manual_seed(seed + 1)
    mel, cache = cfm(mu=mu.clone(), mask=mask, n_timesteps=n_steps, temperature=1.0, spks=spks, cond=cond,
                     prompt_len=prompt, cache=None)
    torch.save(dict(kind="euler", n_blocks=n_blocks, n_mid=n_mid, wseed=WSEED, wsum=wsum(sd), mu=mu, spks=spks,
                    cond=cond, mask=mask, z=z, n_steps=n_steps, prompt=prompt, mel=mel.clone(), cache=cache.clone()),
               os.path.join(HERE, name + ".pt"))
    print(name, "mel sum", float(mel.sum()), "absmax", float(mel.abs().max()), tuple(cache.shape))


def regulator_case(name, B, n_src, ylens, seed, full):
    """The real InterpolateRegulator (modules.py:800-837) on seeded weights: forward, d(sum(out * R))/dx, and the two
    inference layouts. `full` keeps whole tensors, otherwise checksums + a few rows (the 32 x 400 benchmarked shape)."""
    reg = ref_modules.InterpolateRegulator(channels=80, sampling_ratios=(1, 1, 1, 1), out_channels=80, groups=1)
    spec = {k: tuple(v.shape) for k, v in reg.state_dict().items()}
    sd = O.synth_regulator_state_dict(spec, WSEED)
    reg.load_state_dict(sd, strict=True)
    reg.eval()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, n_src, 80, generator=g).requires_grad_(True)
    yl = torch.tensor(ylens)
    R = torch.randn(B, int(yl.max()), 80, generator=g)
    out, _ = reg(x, yl)
    (out * R).sum().backward()
    fx = dict(spec=spec, wseed=WSEED, seed=seed, B=B, n_src=n_src, ylens=list(ylens), wsum=wsum(sd),
              out_sum=csum(out.detach()), dx_sum=csum(x.grad))
    if full:
        fx.update(x=x.detach().clone(), R=R, out=out.detach().clone(), dx=x.grad.clone())
        x1 = torch.randn(1, 30, 80, generator=g)
        x2 = torch.randn(1, 75, 80, generator=g)
        m2 = int(75 / 50 * 22050 / 256)
        with torch.no_grad():
            o1, n1 = reg.inference(x1, x2, 52, m2, 50)
            x3 = torch.randn(1, 33, 80, generator=g)
            o2, n2 = reg.inference(x3[:, :0], x3, 0, 56, 50)
        fx.update(inf_x1=x1, inf_x2=x2, inf_len=(52, m2), inf_out=o1, inf_total=int(n1), inf_short_x=x3, inf_short_len=56,
                  inf_short_out=o2)
    else:
        fx.update(out_rows=out.detach()[:, ::97].clone(), dx_rows=x.grad[:, ::53].clone())
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "out", tuple(out.shape), "|out|", float(out.abs().mean()), "|dx|", float(x.grad.abs().mean()))


def structure_checks():
    """Our module tree == the reference's (names, shapes, and same-seed random init)."""
    for nb, nm in [(1, 1), (4, 12)]:
        ref_utils.set_all_random_seed(4321)
        a = ref_modules.ConditionalDecoder(320, 80, channels=(256, 256), dropout=0.0, attention_head_dim=64,
                                           n_blocks=nb, num_mid_blocks=nm, num_heads=8, act_fn='gelu')
        ref_utils.set_all_random_seed(4321)
        b = our_modules.ConditionalDecoder(320, 80, channels=(256, 256), dropout=0.0, attention_head_dim=64,
                                           n_blocks=nb, num_mid_blocks=nm, num_heads=8, act_fn='gelu')
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys()), "state_dict key order differs"
        for k in sa:
            assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k
        print("structure ok", nb, nm, len(sa), "keys; same-seed init identical")
    spec = {k: tuple(v.shape) for k, v in sa.items()}
    torch.save(dict(keys=list(spec.keys()), shapes=[spec[k] for k in spec]), os.path.join(HERE, "estimator_spec_300m.pt"))


if __name__ == "__main__":
    if sys.argv[1:] == ["inputgrads"]:      # added later: leaves the other fixtures untouched
        input_grad_case("inputgrads_tiny_prompt", "train_tiny_prompt")
        input_grad_case("inputgrads_c1", "train_c1")
        sys.exit(0)
    if sys.argv[1:] == ["dropout"]:
        dropout_case("dropout_tiny_prompt", "train_tiny_prompt", 0.25, 31)
        sys.exit(0)
    if sys.argv[1:] == ["bench"]:
        fml = lambda k: any(s in k for s in ("down_blocks.0.1.0.", "mid_blocks.5.1.2.", "up_blocks.1.1.3."))
        only = os.environ.get("CVFLOW_GOLDEN_ONLY", "")
        if not only or only == "padprompt":
            train_case("train_tiny_padprompt", 1, 1, 3, 64, [64, 40, 33], [50, 30, 25], 101, 9, lambda k: True)
        if not only or only == "dropout":
            dropout_bench_case("dropout_c1", "train_c1", 0.05, 33, fml)
        if not only or only == "euler":
            euler_bench_case("euler_c2", 700, 200, 10, 5)
        if not only or only == "train":
            train_bench_case("train_c3", 32, 400, 99, 7, fml)
        sys.exit(0)
    if sys.argv[1:] == ["regulator"]:
        regulator_case("regulator_tiny", 3, 47, [81, 60, 33], 17, True)
        lens = bench_batch(32, 400, 99)[-1]
        regulator_case("regulator_c3", 32, int(int(lens.max()) * 256 * 50 / 22050), [int(v) for v in lens], 18, False)
        sys.exit(0)
    if sys.argv[1:] == ["flowmodel"]:
        flow_model_case("flowmodel_tiny", 21, 3, 5)
        sys.exit(0)
    if sys.argv[1:] == ["flowinfer"]:
        flow_model_infer_case("flowmodel_infer_tiny", 21, 9)
        sys.exit(0)
    structure_checks()
    first_mid_last = lambda k: any(s in k for s in ("down_blocks.0.1.0.", "mid_blocks.5.1.2.", "up_blocks.1.1.3."))
    train_case("train_tiny", 1, 1, 2, 37, [37, 21], None, 99, 7, lambda k: True)
    train_case("train_tiny_prompt", 1, 1, 3, 64, [64, 50, 33], [12, 0, 9], 100, 8, lambda k: True)
    train_case("train_c1", 4, 12, 2, 200, [200, 160], None, 99, 7, first_mid_last)
    train_case("train_c1_prompt", 4, 12, 2, 200, [200, 160], [30, 0], 99, 7, first_mid_last)
    estimator_case("estimator_tiny", 1, 1, [16, 77, 130, 257], 5)
    estimator_case("estimator_300m", 4, 12, [100, 301], 6)
    euler_case("euler_tiny", 1, 1, 90, 30, 10, 11)
    euler_case("euler_300m", 4, 12, 140, 40, 10, 12)
