"""Two NCCL ranks on two GPUs through the CUDA path (skipped with fewer than 2 devices; run with `gpurun --gpus 2`):
the data-parallel optimiser step (chunked allreduce of the flat LoRA bucket on a side stream behind the backward's
per-chunk events, fused clip + AdamW) must equal AdamW applied to the MEAN of the two single-rank CUDA gradients, and
leave identical parameters on both ranks -- eager, as a captured whole-step graph, and with gradient accumulation as the
micro-step graph + optimiser-tail graph pair."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, graphed, acc, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    tr = None
    try:
        from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
        from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
        from tests.helpers import build_estimator
        est, _, _ = build_estimator(1, 2, lora_r=8)
        # different LoRA values per rank BEFORE the trainer: its constructor must broadcast rank 0's
        if rank == 1:
            with torch.no_grad():
                for n, p in est.named_parameters():
                    if "lora_" in n:
                        p.add_(0.01)
        est = est.to(dev).train()
        cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
        tr = FlowLoRATrainer(cfm, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, accumulate=acc)
        assert len(tr._chunks) >= 2, "gradient chunks not set up"
        p0 = tr.ne.param_bucket.clone()
        both = [torch.empty_like(p0) for _ in range(world)]
        dist.all_gather(both, p0)
        assert torch.equal(both[0], both[1]), "replicas differ after construction"
        g = torch.Generator().manual_seed(100 + rank)          # each rank its own shard
        B, T = 3, 80
        x1, mu = torch.randn(B, 80, T, generator=g).to(dev), torch.randn(B, 80, T, generator=g).to(dev)
        spks, cond = torch.randn(B, 80, generator=g).to(dev), torch.zeros(B, 80, T, device=dev)
        mask = torch.ones(B, 1, T, device=dev)
        mask[1, :, 60 - 7 * rank:] = 0
        # single-rank gradient of this shard (same draws as the step below)
        torch.manual_seed(7 + rank)
        for _ in range(acc):                                  # `acc` micro-batches (fresh draws each) accumulate loss / acc
            loss, _ = cfm.compute_loss(x1, mask, mu, spks, cond=cond)
            (loss / acc).backward()
        # drop the eager autograd graph: it keeps the parameters' AccumulateGrad nodes (created on the default stream)
        # alive, and a later CAPTURED backward that reuses them would make the capturing stream wait on the uncaptured
        # default stream (cudaErrorStreamCaptureIsolation)
        del loss, _
        torch.cuda.synchronize()
        g_local = tr.ne.grad_bucket.clone()
        tr.ne.grad_bucket.zero_()
        gs = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(gs, g_local)
        g_mean = (gs[0] + gs[1]) / world
        # reference update: clip_grad_norm_(1.0) + torch AdamW on the mean gradient
        ref = torch.nn.Parameter(p0.clone())
        ref.grad = g_mean.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=0.01)
        opt.step()
        # the data-parallel step through the trainer
        torch.manual_seed(7 + rank)
        for _ in range(acc):       # accumulate > 1, graphed: the micro-step graph per call, the optimiser-tail graph after the last
            if graphed:
                tr.train_step_graphed(x1, mask, mu, spks, cond)
            else:
                tr.train_step(x1, mask, mu, spks, cond)
        assert tr.step_count == 1
        torch.cuda.synchronize()
        got = tr.ne.param_bucket.clone()
        dist.all_gather(both, got)
        assert torch.equal(both[0], both[1]), "replicas diverged after the step"
        err = float((got - ref.data).abs().max())
        moved = float((ref.data - p0).abs().max())
        assert moved > 0 and err <= 2e-3 * moved + 1e-7, (err, moved)
        out[rank] = (err, moved)
    finally:
        if graphed and tr is not None:
            tr._graph = None
        torch.cuda.synchronize()
        dist.destroy_process_group()


@pytest.mark.parametrize("graphed,acc", [(False, 1), (True, 1), (True, 2)])
def test_two_rank_nccl_step_equals_mean_gradient_adamw(graphed, acc):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 400) + (50 if graphed else 0) + 25 * (acc - 1)
    mp.spawn(_worker, args=(2, port, graphed, acc, out), nprocs=2, join=True)
    assert len(out) == 2
