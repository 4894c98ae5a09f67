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
