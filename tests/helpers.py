"""Shared test helpers: golden loading, synthetic weights, LoRA scaling maps."""
import os

import torch

from oracle import flow_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def build_estimator(n_blocks, n_mid, lora_r=0, lora_alpha=16, wseed=1234, targets=('to_q', 'to_k', 'to_v', 'to_out')):
    """Our ConditionalDecoder with name-seeded synthetic weights (identical to the golden runs)."""
    from cosyvoice_lora_finetune_framework_b200 import lora, modules
    est = modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0,
                                     attention_head_dim=64, n_blocks=n_blocks, num_mid_blocks=n_mid, num_heads=8,
                                     act_fn='gelu')
    stats = None
    if lora_r:
        stats = lora.apply_lora_to_model(est, r=lora_r, lora_alpha=lora_alpha, lora_dropout=0.0,
                                         target_modules=list(targets))
    spec = {k: tuple(v.shape) for k, v in est.state_dict().items()}
    sd = O.synth_state_dict(spec, wseed)
    est.load_state_dict(sd, strict=True)
    return est, sd, stats


def lora_scaling_of(sd, alpha=16):
    return {k[: -len(".lora_A")]: alpha / v.shape[0] for k, v in sd.items() if k.endswith(".lora_A")}


def wsum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))
