"""world_size-2 gloo test of the data-parallel contract: N ranks, each on its own utterance shard
(re-padded to the shard's max length), one allreduce of the flat LoRA-gradient bucket
== the average of N single-process reference runs, one per shard.
The per-rank compute here is the CPU oracle (there is no GPU in this container); the CUDA path
feeds the same bucket layout into the same exchange."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flow_oracle as O
from tests.helpers import build_estimator, lora_scaling_of


def _data():
    g = torch.Generator().manual_seed(3)
    B, T = 4, 40
    lengths = torch.tensor([40, 33, 28, 17])
    batch = dict(x1=torch.randn(B, 80, T, generator=g), mu=torch.randn(B, 80, T, generator=g),
                 spks=torch.randn(B, 80, generator=g), cond=torch.zeros(B, 80, T))
    draws = dict(t_rand=torch.rand(B, 1, 1, generator=g), z=torch.randn(B, 80, T, generator=g), cfg=torch.rand(B, generator=g))
    return batch, lengths, draws


def _shard_grads(sd, rank, world):
    from cosyvoice_lora_finetune_framework_b200.parallel import shard_batch, shard_bounds
    batch, lengths, draws = _data()
    sb, lens = shard_batch({**batch, "z": draws["z"]}, lengths, rank, world)
    lo, hi = shard_bounds(4, rank, world)
    T = sb["x1"].shape[-1]
    mask = (~O.make_pad_mask(lens, T)).float().unsqueeze(1)
    P = {k: v.clone().requires_grad_(k.endswith(("lora_A", "lora_B"))) for k, v in sd.items()}
    loss, _, _ = O.cfm_compute_loss(P, sb["x1"], mask, sb["mu"], sb["spks"], sb["cond"], None, draws["t_rand"][lo:hi],
                                    sb["z"], draws["cfg"][lo:hi], lora_scaling=lora_scaling_of(sd))
    loss.backward()
    names = [k for k in P if P[k].requires_grad]
    return torch.cat([P[k].grad.reshape(-1) for k in names]), float(loss)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from cosyvoice_lora_finetune_framework_b200.parallel import allreduce_mean_
    _, sd, _ = build_estimator(1, 1, lora_r=8)
    bucket, loss = _shard_grads(sd, rank, world)
    allreduce_mean_(bucket)
    if rank == 0:
        ret["bucket"] = bucket.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_mean_of_shard_runs():
    from cosyvoice_lora_finetune_framework_b200.parallel import shard_bounds
    assert [shard_bounds(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_bounds(4, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 4)]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    _, sd, _ = build_estimator(1, 1, lora_r=8)
    g0, _ = _shard_grads(sd, 0, 2)
    g1, _ = _shard_grads(sd, 1, 2)
    want = (g0 + g1) / 2
    assert torch.allclose(ret["bucket"], want, atol=1e-7, rtol=1e-5)
    # and it is NOT the single-run gradient of the concatenated batch (per-shard padding + normaliser)
    gall, _ = _shard_grads(sd, 0, 1)
    assert not torch.allclose(gall, want, atol=1e-6)
