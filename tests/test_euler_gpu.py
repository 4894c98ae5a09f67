"""Euler-ODE inference with CFG (CUDA-graph replayed) vs the reference's golden mel."""
import pytest
import torch

from tests.helpers import build_estimator, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("name", ["euler_tiny", "euler_300m"])
def test_euler_matches_reference(name, graph):
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    fx = load_golden(name)
    est, _, _ = build_estimator(fx["n_blocks"], fx["n_mid"])
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est.cuda().eval()).eval()
    cfm.use_cuda_graph = graph
    c = lambda k: fx[k].cuda()
    for rep in range(2 if graph else 1):          # second call replays the captured graph
        mel, cache = cfm._forward_with_noise(c("z").clone(), c("mu").clone(), c("mask"), fx["n_steps"], c("spks"),
                                             c("cond"), prompt_len=fx["prompt"])
        assert cache.shape == fx["cache"].shape and torch.equal(cache.cpu(), fx["cache"])   # exact: pure indexing
        err = (mel.cpu() - fx["mel"]).abs().max().item()
        # north-star: Euler-solved mel <= 1e-2 max-abs vs the fp32 reference (relative to its range)
        assert err <= 1e-2 * fx["mel"].abs().max().item(), (rep, err, fx["mel"].abs().max().item())
    assert mel.dtype == torch.float32


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("graph", [False, True])
def test_benchmarked_euler_shape(graph, dtype):
    """configs[1], the shape bench.py's inference leg times: T = 700 (200 prompt + 500 target), 10 steps + CFG, 300M
    estimator, vs the real reference's fp32 mel. fp16 operands: <= 1e-2 of the mel range (north-star); bf16 operands:
    within 1.5x the max-abs error the reference itself makes under bf16 autocast (recorded in the fixture)."""
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    from tests.helpers import bench_euler_inputs
    fx = load_golden("euler_c2")
    mu, spks, cond, mask, z = (v.cuda() for v in bench_euler_inputs(fx))
    est, _, _ = build_estimator(fx["n_blocks"], fx["n_mid"])
    est = est.cuda().eval()
    est.cvflow_dtype = dtype
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est).eval()
    cfm.use_cuda_graph = graph
    rng = fx["mel"].abs().max().item()
    for rep in range(2 if graph else 1):
        mel, cache = cfm._forward_with_noise(z.clone(), mu.clone(), mask, fx["n_steps"], spks, cond, prompt_len=fx["prompt"])
        assert torch.equal(cache.cpu(), fx["cache"])
        err = (mel.cpu() - fx["mel"]).abs().max().item()
        print("euler_c2 %s graph=%s rep %d: max-abs %.3e (range %.2f; reference under bf16 autocast %.3e)" %
              (dtype, graph, rep, err, rng, fx["ref_autocast"]["bf16"]["max_abs"]))
        if dtype == torch.float16:
            assert err <= 1e-2 * rng, (err, rng)
        else:
            assert err <= 1.5 * fx["ref_autocast"]["bf16"]["max_abs"], (err, fx["ref_autocast"]["bf16"])


def test_public_forward_draws_noise_like_reference():
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    est, _, _ = build_estimator(1, 1)
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est.cuda().eval()).eval()
    mu = torch.randn(1, 80, 50, device="cuda")
    spks = torch.randn(1, 80, device="cuda")
    torch.manual_seed(3)
    z = torch.randn_like(mu)
    torch.manual_seed(3)
    mel, cache = cfm(mu=mu.clone(), mask=torch.ones(1, 1, 50, device="cuda"), n_timesteps=4, spks=spks,
                     cond=torch.zeros(1, 80, 50, device="cuda"), prompt_len=0, cache=None)
    assert torch.equal(cache[:, :, :, 0], z[:, :, -34:])
    assert mel.shape == (1, 80, 50) and torch.isfinite(mel).all()
