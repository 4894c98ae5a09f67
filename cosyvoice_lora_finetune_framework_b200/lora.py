"""LoRA inject / save / load / merge with the reference's semantics and checkpoint key layout
(reference cosyvoice_flow_finetune/lora.py).

Parity notes reproduced on purpose (SURVEY.md "parity traps" 2-4):
  * injection matches `target in child_name` on the *immediate* child name, so `to_out`
    (a ModuleList whose Linear child is named "0") is never wrapped (lora.py:178-182);
  * lora_B is N(0, 0.01), not zero; lora_A is kaiming-uniform(a=sqrt(5)) (lora.py:57-62);
  * `merge_lora_weights` adds into the frozen weight in place and is not idempotent; the merged
    state-dict lists LoRA-wrapped layers first, then the other parameters, then buffers
    (lora.py:259-323).

`LoRALinear.forward` is only used outside the fused estimator (e.g. an encoder Linear); inside
`ConditionalDecoder` the q/k/v LoRA layers are consumed by the CUDA path as parameter holders.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class LoRALinear(nn.Module):
    """y = W x (+ b) + (alpha / r) * B (A drop(x)); W, b frozen."""

    def __init__(self, original_layer: nn.Linear, r: int = 8, lora_alpha: int = 16,
                 lora_dropout: float = 0.1):
        super().__init__()
        self.original_layer = original_layer
        self.r = r
        self.lora_alpha = lora_alpha
        self.scaling = lora_alpha / r
        for p in original_layer.parameters():
            p.requires_grad = False
        self.lora_A = nn.Parameter(torch.zeros(r, original_layer.in_features))
        self.lora_B = nn.Parameter(torch.zeros(original_layer.out_features, r))
        self.lora_dropout = nn.Dropout(p=lora_dropout) if lora_dropout > 0 else nn.Identity()
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        nn.init.normal_(self.lora_B, mean=0.0, std=0.01)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        base = self.original_layer(x)
        low = F.linear(self.lora_dropout(x), self.lora_A.to(x.dtype))
        low = F.linear(low, self.lora_B.to(x.dtype))
        return base + low * self.scaling


class LoRAConv1d(nn.Module):
    """LoRA on a 1x1 Conv1d: two bias-free 1x1 convs (r channels in between)."""

    def __init__(self, original_layer: nn.Conv1d, r: int = 8, lora_alpha: int = 16,
                 lora_dropout: float = 0.1):
        super().__init__()
        self.original_layer = original_layer
        self.r = r
        self.lora_alpha = lora_alpha
        self.scaling = lora_alpha / r
        for p in original_layer.parameters():
            p.requires_grad = False
        self.lora_A = nn.Conv1d(original_layer.in_channels, r, kernel_size=1, bias=False)
        self.lora_B = nn.Conv1d(r, original_layer.out_channels, kernel_size=1, bias=False)
        self.lora_dropout = nn.Dropout(p=lora_dropout) if lora_dropout > 0 else nn.Identity()
        nn.init.kaiming_uniform_(self.lora_A.weight, a=math.sqrt(5))
        nn.init.normal_(self.lora_B.weight, mean=0.0, std=0.01)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        base = self.original_layer(x)
        low = F.conv1d(self.lora_dropout(x), self.lora_A.weight.to(x.dtype))
        low = F.conv1d(low, self.lora_B.weight.to(x.dtype))
        return base + low * self.scaling


DEFAULT_TARGETS = ['to_q', 'to_k', 'to_v', 'linear_q', 'linear_k', 'linear_v', 'linear_out',
                   'w_1', 'w_2', 'linear_pos']


def apply_lora_to_model(model: nn.Module, r: int = 8, lora_alpha: int = 16, lora_dropout: float = 0.1,
                        target_modules: Optional[List[str]] = None) -> Dict[str, int]:
    """Wrap matching Linear / 1x1-Conv1d children with LoRA, freeze everything else, return stats."""
    targets = set(DEFAULT_TARGETS if target_modules is None else target_modules)
    original_params = sum(p.numel() for p in model.parameters())
    stats = {'layers': 0, 'params': 0}

    def visit(parent: nn.Module):
        for child_name, child in list(parent.named_children()):
            if any(t in child_name for t in targets):
                wrapped = None
                if isinstance(child, nn.Linear):
                    wrapped = LoRALinear(child, r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout)
                    added = wrapped.lora_A.numel() + wrapped.lora_B.numel()
                elif isinstance(child, nn.Conv1d) and child.kernel_size[0] == 1:
                    wrapped = LoRAConv1d(child, r=r, lora_alpha=lora_alpha, lora_dropout=lora_dropout)
                    added = wrapped.lora_A.weight.numel() + wrapped.lora_B.weight.numel()
                if wrapped is not None:
                    setattr(parent, child_name, wrapped)
                    stats['layers'] += 1
                    stats['params'] += added
            # the walk descends into the *old* child, exactly like the reference (lora.py:209)
            visit(child)

    visit(model)
    for name, p in model.named_parameters():
        if 'lora_' not in name:
            p.requires_grad = False
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if hasattr(model, '_cvflow_invalidate'):
        model._cvflow_invalidate()
    return {
        'replaced_layers': stats['layers'],
        'original_params': original_params,
        'lora_params': stats['params'],
        'trainable_params': trainable,
        'trainable_ratio': trainable / original_params * 100,
    }


def get_lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    return {n: p.data.clone() for n, p in model.named_parameters() if 'lora_' in n}


def save_lora_weights(model: nn.Module, path: str):
    sd = get_lora_state_dict(model)
    torch.save(sd, path)
    print(f"Saved LoRA weights: {len(sd)} tensors to {path}")


def load_lora_weights(model: nn.Module, path: str):
    sd = torch.load(path, map_location='cpu')
    own = model.state_dict()
    for name, value in sd.items():
        if name in own:
            own[name].copy_(value)
    print(f"Loaded LoRA weights: {len(sd)} tensors from {path}")


def merge_lora_weights(model: nn.Module):
    """W += (B @ A) * scaling, in place (calling it twice adds twice, as in the reference)."""
    with torch.no_grad():
        for _, m in model.named_modules():
            if isinstance(m, LoRALinear):
                m.original_layer.weight.add_(m.lora_B @ m.lora_A * m.scaling)
            elif isinstance(m, LoRAConv1d):
                delta = torch.einsum('ori,ric->oic', m.lora_B.weight, m.lora_A.weight) * m.scaling
                m.original_layer.weight.add_(delta)
    if hasattr(model, '_cvflow_invalidate'):
        model._cvflow_invalidate()
    print("LoRA weights merged into original model")


def get_merged_state_dict(model: nn.Module) -> dict:
    """Merge, then emit a state-dict in the un-LoRA'd key layout (loadable by the stock model)."""
    merge_lora_weights(model)
    out = {}
    for name, m in model.named_modules():
        if isinstance(m, (LoRALinear, LoRAConv1d)):
            out[f"{name}.weight"] = m.original_layer.weight.data.clone()
            if m.original_layer.bias is not None:
                out[f"{name}.bias"] = m.original_layer.bias.data.clone()
    for name, p in model.named_parameters():
        if 'lora_A' in name or 'lora_B' in name or 'original_layer' in name:
            continue
        out[name] = p.data.clone()
    for name, buf in model.named_buffers():
        if 'lora_' not in name and 'original_layer' not in name:
            out[name] = buf.clone()
    print(f"Exported merged state_dict with {len(out)} keys")
    return out
