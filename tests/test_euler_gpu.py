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


def test_time_embed_entry_point_vs_torch():
    """cvflow_time_embed (SinusoidalPosEmb scale 1000 -> Linear -> SiLU -> Linear, modules.py:27-57) stand-alone."""
    import torch.nn.functional as F
    from oracle import flow_oracle as O
    from cosyvoice_lora_finetune_framework_b200 import _estimator as E
    est, sd, _ = build_estimator(1, 1)
    ne = E.native_of(est.cuda().eval())
    t = torch.tensor([0.0, 1e-3, 0.25, 0.5, 0.9999, 1.0, 0.3333])
    emb = O.sinusoidal_embedding(t)
    ref = F.linear(F.silu(F.linear(emb, sd["time_mlp.linear_1.weight"], sd["time_mlp.linear_1.bias"])),
                   sd["time_mlp.linear_2.weight"], sd["time_mlp.linear_2.bias"])
    got = ne.time_embed(t.cuda()).cpu()
    assert got.shape == (7, 1024)
    assert torch.allclose(got, ref, atol=2e-4 * float(ref.abs().max()) + 1e-5, rtol=1e-3), float((got - ref).abs().max())


def test_solve_capture_replay_c_abi():
    """cvflow_solve_capture / cvflow_solve_replay called directly (no ConditionalCFM): the library-owned graph reproduces
    the eager per-step launches bit for bit, can be replayed on new inputs, and a second (T, n_steps) coexists."""
    import ctypes as C
    from cosyvoice_lora_finetune_framework_b200 import _estimator as E, _native as N
    est, _, _ = build_estimator(1, 1)
    ne = E.native_of(est.cuda().eval())
    L = E._lib()
    dev = torch.device("cuda")

    def make(T, n, seed):
        g = torch.Generator().manual_seed(seed)
        return dict(x=torch.randn(1, 80, T, generator=g).to(dev), mu=torch.randn(1, 80, T, generator=g).to(dev),
                    spks=torch.randn(1, 80, generator=g).to(dev), cond=torch.zeros(1, 80, T, device=dev),
                    mask=torch.ones(1, T, device=dev), t=torch.linspace(0, 0.9, n).to(dev), dt=torch.full((n,), 1.0 / n).to(dev),
                    d=torch.empty(2, 80, T, device=dev), keep=torch.tensor([1.0, 0.0], device=dev))

    def eager(st, T, n):
        x = st["x"].clone()
        for k in range(n):
            ne.forward(x, st["mask"], st["mu"], st["t"][k:k + 1], st["spks"], st["cond"], keep=st["keep"], iso_len=0,
                       training=False, B=2, out=st["d"])
            N.check(L.cvflow_euler_update(x.data_ptr(), st["d"].data_ptr(), st["dt"].data_ptr(), k, 0.7, 80 * T, E._stream()), "eu")
        return x

    cap = torch.cuda.Stream()
    sts = {}
    for (T, n, seed) in [(96, 4, 1), (57, 3, 2)]:
        st = make(T, n, seed)
        want = eager(st, T, n)
        ne._workspace(2, T, False)
        if "ws" in sts:
            assert sts["ws"] == ne.ws.data_ptr(), "the arena must not move between captures in this test"
        sts["ws"] = ne.ws.data_ptr()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            N.check(L.cvflow_solve_capture(ne.handle, T, n, 0.7, st["x"].data_ptr(), st["mask"].data_ptr(), st["mu"].data_ptr(),
                                           st["spks"].data_ptr(), st["cond"].data_ptr(), st["t"].data_ptr(), st["dt"].data_ptr(),
                                           st["d"].data_ptr(), C.c_void_p(cap.cuda_stream)), "cvflow_solve_capture")
        torch.cuda.current_stream().wait_stream(cap)
        sts[(T, n)] = (st, want)
    for (T, n), (st, want) in [(k, v) for k, v in sts.items() if k != "ws"]:
        x0 = st["x"].clone()
        N.check(L.cvflow_solve_replay(ne.handle, T, n, E._stream()), "cvflow_solve_replay")
        torch.cuda.synchronize()
        assert torch.equal(st["x"], want), (T, n)
        # new utterance through the same graph
        st["x"].copy_(x0 * 0.5)
        st["mu"].mul_(-1.0)
        want2 = eager(st, T, n)
        N.check(L.cvflow_solve_replay(ne.handle, T, n, E._stream()), "cvflow_solve_replay")
        torch.cuda.synchronize()
        assert torch.equal(st["x"], want2)
    assert L.cvflow_solve_replay(ne.handle, 31, 2, E._stream()) != 0          # never captured: a loud error, no launch
    assert b"no captured solve" in L.cvflow_last_error()
    N.check(L.cvflow_solve_release(ne.handle), "cvflow_solve_release")
    assert L.cvflow_solve_replay(ne.handle, 96, 4, E._stream()) != 0
