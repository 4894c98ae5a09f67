"""Data-parallel plumbing of the flow-LoRA fine-tune: utterance sharding and the single gradient
exchange. Backend-agnostic (NCCL on the B200 box, gloo in the CPU tests)."""
from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], lengths: torch.Tensor, rank: int, world: int):
    """Take this rank's utterances and re-pad them to the shard's own max length -- what `world`
    independent reference data loaders would produce (GroupNorm sees the padding, so the padded
    length is part of the numerics; SURVEY.md trap 1). Tensors with a trailing time axis are cropped."""
    lo, hi = shard_bounds(int(lengths.shape[0]), rank, world)
    lens = lengths[lo:hi]
    t_max = int(lens.max()) if hi > lo else 0
    full_t = max(v.shape[-1] for v in batch.values() if v.dim() == 3)
    out = {}
    for k, v in batch.items():
        s = v[lo:hi]
        if s.dim() == 3 and s.shape[-1] == full_t:
            s = s[..., :t_max]
        out[k] = s.contiguous()
    return out, lens


def allreduce_mean_(bucket: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """In-place mean over ranks of the flat fp32 LoRA-gradient bucket (== DDP gradient averaging)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        bucket.div_(dist.get_world_size(group))
    return bucket
