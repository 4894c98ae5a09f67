"""Mask / seed helpers of the flow path (API mirror of the reference's utils.py).

Integer work is exact: `make_pad_mask` and `mask_to_bias` reproduce the reference bit for bit
(reference cosyvoice_flow_finetune/utils.py:12-41,103-109). Inside the CUDA estimator the
[B, L, L] additive bias is never materialised - the attention kernel derives it from the
[B, L] key mask - so `mask_to_bias` exists for API compatibility and for the oracle tests.
"""
import random

import numpy as np
import torch


def set_all_random_seed(seed):
    """Seed python, numpy and torch (CPU + every CUDA device), as reference utils.py:12-17."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def make_pad_mask(lengths: torch.Tensor, max_len: int = 0) -> torch.Tensor:
    """True where position >= length (i.e. padding). lengths (B,) -> bool (B, max_len).

    >>> make_pad_mask(torch.tensor([5, 3, 2])).int().tolist()
    [[0, 0, 0, 0, 0], [0, 0, 0, 1, 1], [0, 0, 1, 1, 1]]
    """
    if max_len <= 0:
        max_len = int(lengths.max().item())
    positions = torch.arange(max_len, dtype=torch.int64, device=lengths.device)
    return positions[None, :] >= lengths[:, None]


def mask_to_bias(mask: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """bool keep-mask -> additive attention bias: 0 where True, -1e10 where False."""
    if mask.dtype != torch.bool:
        raise AssertionError("mask_to_bias expects a bool mask")
    if dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise AssertionError("unsupported bias dtype %s" % dtype)
    keep = mask.to(dtype)
    return (1.0 - keep) * -1.0e10


def pad_list(xs, pad_value):
    """Stack variable-length tensors (1-3 dims) into one batch filled with pad_value."""
    longest = max(int(x.shape[0]) for x in xs)
    tail = tuple(xs[0].shape[1:])
    if len(tail) > 2:
        raise ValueError("Unsupported ndim: %d" % xs[0].ndim)
    out = xs[0].new_full((len(xs), longest) + tail, pad_value)
    for row, x in zip(out, xs):
        row[: x.shape[0]] = x
    return out
