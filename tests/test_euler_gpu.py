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
