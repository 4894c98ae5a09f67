"""Shared test helpers: golden loading, synthetic weights, LoRA scaling maps."""
import os

import torch

from oracle import flow_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def build_estimator(n_blocks, n_mid, lora_r=0, lora_alpha=16, wseed=1234, targets=('to_q', 'to_k', 'to_v', 'to_out'),
                    lora_dropout=0.0):
    """Our ConditionalDecoder with name-seeded synthetic weights (identical to the golden runs)."""
    from cosyvoice_lora_finetune_framework_b200 import lora, modules
    est = modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0,
                                     attention_head_dim=64, n_blocks=n_blocks, num_mid_blocks=n_mid, num_heads=8,
                                     act_fn='gelu')
    stats = None
    if lora_r:
        stats = lora.apply_lora_to_model(est, r=lora_r, lora_alpha=lora_alpha, lora_dropout=lora_dropout,
                                         target_modules=list(targets))
    spec = {k: tuple(v.shape) for k, v in est.state_dict().items()}
    sd = O.synth_state_dict(spec, wseed)
    est.load_state_dict(sd, strict=True)
    return est, sd, stats


def lora_scaling_of(sd, alpha=16):
    return {k[: -len(".lora_A")]: alpha / v.shape[0] for k, v in sd.items() if k.endswith(".lora_A")}


def wsum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def dropout_masks(n_tbs, rows, p, seed):
    """Explicit LoRA-dropout keep masks [n_tbs][3][rows][256] (uint8), the same draw as tests/golden/make_golden.py."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n_tbs, 3, rows, 256, generator=g) >= p).to(torch.uint8)


def attention_block_prefixes(n_blocks, n_mid):
    """State-dict prefixes of the transformer blocks in execution order (= block index of the CUDA path)."""
    stages = ["down_blocks.0", "down_blocks.1"] + ["mid_blocks.%d" % i for i in range(n_mid)] + ["up_blocks.0", "up_blocks.1"]
    return ["%s.1.%d" % (s, j) for s in stages for j in range(n_blocks)]


def oracle_dropout_entries(keep, p, prefixes, B, T):
    """`<prefix>.attn1.to_{q,k,v}.lora_dropout_mask` entries for the oracle's parameter dict: block i at length L uses
    rows [0, B*L) of keep[i, proj] viewed as [B, L, 256]."""
    out = {}
    T2 = (T + 1) // 2
    for i, q in enumerate(prefixes):
        L = T if (q.startswith("down_blocks.0") or q.startswith("up_blocks.1")) else T2
        for j, pn in enumerate(("to_q", "to_k", "to_v")):
            out["%s.attn1.%s.lora_dropout_mask" % (q, pn)] = keep[i, j, : B * L].view(B, L, 256).float() / (1.0 - p)
    return out


# ------------------------------------------------------------------------------------------------------
# Fixtures at the benchmarked shapes (train_c3, euler_c2) do not store their big inputs: they are regenerated
# from the seeds with the same torch.Generator call sequence as bench.py / make_golden.py and verified against
# the checksums the fixture recorded from the reference run.
# ------------------------------------------------------------------------------------------------------
def csum(t):
    return [float(t.double().sum()), float(t.double().abs().sum())]


def check_sums(tensors, checks):
    for k, t in tensors.items():
        s = csum(t)
        assert abs(s[0] - checks[k][0]) <= 1e-6 * (1.0 + abs(checks[k][1])) and \
            abs(s[1] - checks[k][1]) <= 1e-6 * (1.0 + abs(checks[k][1])), ("regenerated input differs from the fixture", k)


def bench_train_inputs(fx):
    """(x1, mask, mu, spks, cond, t_rand, z, cfg_rand) of a `train_bench` fixture: bench.py::make_batch + the three
    draws of compute_loss under torch.manual_seed(step_seed)."""
    import bench
    B, T = fx["B"], fx["T"]
    batch, lens = bench.make_batch(B, T, fx["batch_seed"], "cpu")
    assert torch.equal(lens, fx["lengths"])
    torch.manual_seed(fx["step_seed"])
    t_rand = torch.rand([B, 1, 1])
    z = torch.randn(B, 80, T)
    cfg_rand = torch.rand(B)
    check_sums(dict(x1=batch["x1"], mu=batch["mu"], spks=batch["spks"], z=z, t_rand=t_rand, cfg_rand=cfg_rand), fx["checks"])
    return batch["x1"], batch["mask"], batch["mu"], batch["spks"], batch["cond"], t_rand, z, cfg_rand


def bench_euler_inputs(fx):
    """(mu, spks, cond, mask, z) of an `euler_bench` fixture (bench.py's inference leg, seed 5)."""
    T, P = fx["T"], fx["prompt"]
    g = torch.Generator().manual_seed(fx["seed"])
    mu = torch.randn(1, 80, T, generator=g)
    spks = torch.randn(1, 80, generator=g)
    cond = torch.zeros(1, 80, T)
    cond[:, :, :P] = torch.randn(1, 80, P, generator=g)
    torch.manual_seed(fx["seed"] + 1)
    z = torch.randn_like(mu)
    check_sums(dict(mu=mu, spks=spks, cond=cond, z=z), fx["checks"])
    return mu, spks, cond, torch.ones(1, 1, T), z


def grad_errors(grads, ref):
    """bucket rel-L2 and per-tensor rel-L2 (median, max) of `grads` against `ref` over the tensors in `ref`."""
    num = sum((grads[k].cpu() - ref[k]).double().pow(2).sum() for k in ref)
    den = sum(ref[k].double().pow(2).sum() for k in ref)
    per = sorted(float((grads[k].cpu() - ref[k]).double().norm() / (ref[k].double().norm() + 1e-30)) for k in ref)
    return dict(bucket_rel_l2=float((num / den).sqrt()), tensor_rel_l2_median=per[len(per) // 2], tensor_rel_l2_max=per[-1])


# ------------------------------------------------------------------------------------------------------
# length regulator fixtures (regulator_tiny / regulator_c3): inputs regenerated from the seed, weights from the name hash
# ------------------------------------------------------------------------------------------------------
def regulator_inputs(fx):
    """(state dict, x, ylens, R) of a `regulator` fixture, regenerated exactly as tests/golden/make_golden.py drew them."""
    sd = O.synth_regulator_state_dict(fx["spec"], fx["wseed"])
    assert abs(wsum(sd) - fx["wsum"]) <= 1e-6 * abs(fx["wsum"])
    g = torch.Generator().manual_seed(fx["seed"])
    x = torch.randn(fx["B"], fx["n_src"], 80, generator=g)
    yl = torch.tensor(fx["ylens"])
    R = torch.randn(fx["B"], int(yl.max()), 80, generator=g)
    return sd, x, yl, R


def build_regulator(sd):
    from cosyvoice_lora_finetune_framework_b200.encoder import InterpolateRegulator
    reg = InterpolateRegulator(channels=80, sampling_ratios=(1, 1, 1, 1), out_channels=80, groups=1)
    reg.load_state_dict(sd, strict=True)
    for p in reg.parameters():
        p.requires_grad_(False)
    return reg


def close_sums(t, ref, tol):
    s = csum(t)
    return abs(s[0] - ref[0]) <= tol * (1.0 + abs(ref[1])) and abs(s[1] - ref[1]) <= tol * (1.0 + abs(ref[1]))
