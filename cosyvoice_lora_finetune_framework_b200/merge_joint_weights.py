"""Merge LoRA weights out of a joint-training checkpoint into stock-layout `llm.pt` / `flow.pt`
files (API mirror of the reference's merge_joint_weights.py: same functions, flags and output
layout; Lightning prefixes `model.llm.` / `model.flow.` are stripped, :98,155).

  python -m cosyvoice_lora_finetune_framework_b200.merge_joint_weights [--ckpt X] [--llm-only|--flow-only]
         [--llm-output P] [--flow-output P]
"""
import argparse
import os
from typing import Optional

import torch

from .config import JOINT_TRAINING_CONFIG, OUTPUT_DIR, PRETRAINED_MODEL_DIR


def find_latest_joint_checkpoint(output_dir: str, mode: Optional[str] = None) -> Optional[str]:
    """Newest *.ckpt in output_dir, filtered by mode (or preferring 'joint_joint' when mode is None)."""
    names = [f for f in os.listdir(output_dir) if f.endswith('.ckpt')]
    if mode:
        names = [f for f in names if f'joint_{mode}' in f or f'{mode}' in f]
    else:
        preferred = [f for f in names if 'joint_joint' in f]
        names = preferred or names
    if not names:
        return None
    return max((os.path.join(output_dir, f) for f in names), key=os.path.getmtime)


def _load_half(module, state_dict, prefixes):
    own = module.state_dict()
    loaded = 0
    for key, value in state_dict.items():
        for p in prefixes:
            if key.startswith(p):
                key = key[len(p):]
                break
        if key in own and own[key].shape == value.shape:
            own[key] = value
            loaded += 1
    module.load_state_dict(own)
    return loaded


def _merge(ckpt_path, mode, which, output_path):
    from .llm_flow_model import build_joint_model
    from .lora import get_merged_state_dict
    ckpt = torch.load(ckpt_path, map_location='cpu')
    state = ckpt.get('state_dict', ckpt)
    model = build_joint_model(pretrained_path=PRETRAINED_MODEL_DIR, device='cpu', training_mode=mode,
                              llm_lora_config=JOINT_TRAINING_CONFIG.get('llm_lora') if which == 'llm' else None,
                              flow_lora_config=JOINT_TRAINING_CONFIG.get('flow_lora') if which == 'flow' else None)
    half = getattr(model, which)
    n = _load_half(half, state, [f'model.{which}.', f'{which}.'])
    print(f"loaded {n} {which} tensors from {ckpt_path}")
    merged = get_merged_state_dict(half)
    torch.save(merged, output_path)
    print(f"{which} weights saved: {output_path} ({os.path.getsize(output_path) / 1024 / 1024:.1f} MB)")
    return merged


def merge_llm_from_checkpoint(ckpt_path: str, output_path: str):
    return _merge(ckpt_path, 'llm_only', 'llm', output_path)


def merge_flow_from_checkpoint(ckpt_path: str, output_path: str):
    return _merge(ckpt_path, 'flow_only', 'flow', output_path)


def merge_both_from_checkpoint(ckpt_path: str, llm_output: str, flow_output: str):
    """The merge mutates the model in place, so each half is merged on a freshly built model
    (the reference rebuilds before the flow merge for the same reason, :244-264)."""
    return {'llm': merge_llm_from_checkpoint(ckpt_path, llm_output),
            'flow': merge_flow_from_checkpoint(ckpt_path, flow_output)}


def main():
    ap = argparse.ArgumentParser(description="merge joint-training LoRA weights")
    ap.add_argument('--ckpt', type=str, default=None)
    ap.add_argument('--llm-only', action='store_true')
    ap.add_argument('--flow-only', action='store_true')
    ap.add_argument('--llm-output', type=str, default=os.path.join(OUTPUT_DIR, 'llm_merged.pt'))
    ap.add_argument('--flow-output', type=str, default=os.path.join(OUTPUT_DIR, 'flow_merged.pt'))
    a = ap.parse_args()
    ckpt = a.ckpt or find_latest_joint_checkpoint(OUTPUT_DIR)
    if not ckpt or not os.path.exists(ckpt):
        raise SystemExit("no checkpoint found (use --ckpt)")
    if a.llm_only:
        merge_llm_from_checkpoint(ckpt, a.llm_output)
    elif a.flow_only:
        merge_flow_from_checkpoint(ckpt, a.flow_output)
    else:
        merge_both_from_checkpoint(ckpt, a.llm_output, a.flow_output)


if __name__ == '__main__':
    main()
