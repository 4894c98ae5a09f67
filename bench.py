#!/usr/bin/env python
"""Benchmark of the flow-LoRA hot path (BASELINE.json metric: LoRA flow train mel-frames/sec; Euler-ODE
inference RTF reported alongside).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A "step" is one full optimiser step of the data-parallel LoRA fine-tune on one synthetic batch per
rank: CFM interpolation -> estimator forward -> masked loss -> estimator backward -> (N>1) NCCL
allreduce of the flat LoRA-gradient bucket -> fused clip + AdamW -> W_eff refresh. Workload =
BASELINE.json configs[2] (CosyVoice-300M flow estimator, LoRA r=8 on attn1 q/k/v with the reference's
default lora_dropout = 0.05 (config.py:207-216), batch 32 x 400 frames per GPU, weak scaling;
--global-batch G gives strong scaling). Prints ONE JSON line on rank 0.

What the line carries (N = 1): value (device-resident inputs), e2e (pinned host inputs copied every step on a copy
stream, loss read back every step), roofline of the dominant kernel class (GEMM engine) + per-class entries for
attention and the HBM-bound norm passes (CUDA events around every launch of an eager step), cpu_baseline (the
oracle port of the reference on the host cores, same shape), and informational legs: the folded lora_dropout = 0
step, configs[4] (16 x 1500 ragged), the reference algorithm in eager PyTorch bf16 autocast on the same GPU, and
configs[1] inference (10-step CFG Euler solve, T = 700) with its own CPU baseline and roofline.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "flow_lora_train_mel_frames_per_sec"
UNIT = "mel-frames/s"
LORA_DROPOUT = 0.05      # the reference's default (config.py:207-216); SURVEY 8(d): 0.05 for throughput


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cvflow", choices=["cvflow", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0, help="strong scaling: total utterances, split over the ranks")
    ap.add_argument("--frames", type=int, default=400, help="padded mel frames per utterance")
    ap.add_argument("--min-len", type=float, default=0.6, help="ragged lengths are drawn from (min_len * T, T]")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--lora-dropout", type=float, default=LORA_DROPOUT)
    ap.add_argument("--streams", type=int, default=1, help="concurrent batch shards (CUDA streams) per GPU (dropout 0 only)")
    ap.add_argument("--no-inference", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the folded / configs[4] / eager-GPU legs")
    return ap.parse_args()


def per_rank_batch(a, world):
    if a.global_batch:
        if a.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of ranks")
        return a.global_batch // world
    return a.batch


def workload_config(a, world):
    B = per_rank_batch(a, world)
    return {"workload": "configs[2]: CosyVoice-300M flow estimator (16 resnets, 64 transformer blocks), LoRA r=8 "
                        "alpha=16 lora_dropout=%g on attn1 to_q/to_k/to_v, batch %d x %d frames per GPU, ragged lengths in "
                        "(%gT, T]" % (a.lora_dropout, B, a.frames, a.min_len),
            "global_batch": B * world, "frames": a.frames, "parallelism": "dp%d" % world,
            "launch": "whole optimiser step replayed as one CUDA graph (PDL edges between kernels)"
                      + ("; batch shards on %d concurrent streams" % a.streams if a.streams > 1 else ""),
            "l2": "per-step working set (stashed activations ~5 GB) >> 126 MB L2, no explicit flush needed"}


def make_batch(B, T, seed, device, min_len=0.6):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(B, 80, T, generator=g)
    mu = torch.randn(B, 80, T, generator=g)
    spks = torch.randn(B, 80, generator=g)
    cond = torch.zeros(B, 80, T)
    lens = torch.randint(int(min_len * T) + 1, T + 1, (B,), generator=g)
    lens[0] = T
    mask = (torch.arange(T)[None, :] < lens[:, None]).float().unsqueeze(1)
    t = dict(x1=x1, mu=mu, spks=spks, cond=cond, mask=mask)
    return {k: v.to(device) for k, v in t.items()}, lens


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_model(device, dtype, lora_dropout):
    from cosyvoice_lora_finetune_framework_b200 import lora, modules, utils
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    utils.set_all_random_seed(1234)
    est = modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0,
                                     attention_head_dim=64, n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn='gelu')
    stats = lora.apply_lora_to_model(est, r=8, lora_alpha=16, lora_dropout=lora_dropout,
                                     target_modules=['to_q', 'to_k', 'to_v', 'to_out'])
    est = est.to(device).train()
    est.cvflow_dtype = dtype
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, sigma_min=1e-6, t_scheduler='cosine',
                         training_cfg_rate=0.2, inference_cfg_rate=0.7, estimator=est)
    return cfm, est, stats


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1435.3), d.get("hbm_gbs", 6452.2), "measured (MEASURED_PEAKS.json; bf16 sustained, HBM copy)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_gemm_traffic():
    """dram bytes per GEMM launch from the committed `ncu --set full` capture of the step's GEMM shapes
    (profiles/r02_gemm_shape_table.json, launch-count weighted over the shapes of one step); None when absent."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_shape_table.json")
    try:
        d = json.load(open(p))
        return d["traffic_bytes_per_launch_step_weighted"], "profiles/r02_gemm_shape_table.json (ncu --set full, one launch per shape)"
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------------
# the reference algorithm (oracle port: plain PyTorch fp32 + autograd + AdamW), on the host cores or,
# informationally, in eager PyTorch on the GPU ("the existing Blackwell path")
# ---------------------------------------------------------------------------------------------------
def oracle_step_fn(B, T, device, lora_dropout, min_len=0.6, autocast=None):
    from oracle import flow_oracle as O
    from cosyvoice_lora_finetune_framework_b200 import lora, modules, utils
    utils.set_all_random_seed(1234)
    est = modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0,
                                     attention_head_dim=64, n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn='gelu')
    lora.apply_lora_to_model(est, r=8, lora_alpha=16, lora_dropout=0.0, target_modules=['to_q', 'to_k', 'to_v'])
    P = {k: v.detach().clone().to(device) for k, v in est.state_dict().items()}
    train = [k for k in P if k.endswith(("lora_A", "lora_B"))]
    for k in train:
        P[k].requires_grad_(True)
    scaling = {k[:-len(".lora_A")]: 2.0 for k in P if k.endswith(".lora_A")}
    if lora_dropout > 0:
        P["__lora_dropout_p__"] = float(lora_dropout)      # every LoRA layer draws its own mask, like nn.Dropout
    opt = torch.optim.AdamW([P[k] for k in train], lr=1e-4, weight_decay=0.01)
    batch, _ = make_batch(B, T, 99, device, min_len)
    g = torch.Generator(device=device).manual_seed(7)

    def step():
        t_rand = torch.rand(B, 1, 1, generator=g, device=device)
        z = torch.randn(B, 80, T, generator=g, device=device)
        cfg = torch.rand(B, generator=g, device=device)
        if autocast is not None:
            with torch.autocast(device_type=torch.device(device).type, dtype=autocast):
                loss, _, _ = O.cfm_compute_loss(P, batch["x1"], batch["mask"], batch["mu"], batch["spks"], batch["cond"],
                                                None, t_rand, z, cfg, lora_scaling=scaling)
        else:
            loss, _, _ = O.cfm_compute_loss(P, batch["x1"], batch["mask"], batch["mu"], batch["spks"], batch["cond"], None,
                                            t_rand, z, cfg, lora_scaling=scaling)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([P[k] for k in train], 1.0)
        opt.step()
        return float(loss)

    return step


def time_host(fn, n):
    out = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t0)
    return out


def oracle_euler_fn(T, P_, n_steps, device):
    from oracle import flow_oracle as O
    from cosyvoice_lora_finetune_framework_b200 import modules, utils
    utils.set_all_random_seed(1234)
    est = modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0,
                                     attention_head_dim=64, n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn='gelu')
    sd = {k: v.detach().clone().to(device) for k, v in est.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    mu = torch.randn(1, 80, T, generator=g).to(device)
    spk = torch.randn(1, 80, generator=g).to(device)
    cond = torch.zeros(1, 80, T, device=device)
    cond[:, :, :P_] = torch.randn(1, 80, P_, generator=g).to(device)
    mask = torch.ones(1, 1, T, device=device)
    z = torch.randn(1, 80, T, generator=g).to(device)

    def run():
        with torch.no_grad():
            O.cfm_forward(sd, mu.clone(), mask, n_steps, z.clone(), spk, cond, prompt_len=P_)

    return run


def flow_model_batch(B, T, seed, min_len=0.6):
    """A batch in the reference's collate schema (train_joint.py dataset: tokens at 50 Hz, mel at 22050/256 Hz)."""
    g = torch.Generator().manual_seed(seed)
    feat_len = torch.randint(int(min_len * T) + 1, T + 1, (B,), generator=g)
    feat_len[0] = T
    tok_len = (feat_len.float() * 256 * 50 / 22050).long().clamp(min=1)
    N = int(tok_len.max())
    feat = torch.full((B, T, 80), -11.5)
    tok = torch.zeros(B, N, dtype=torch.long)
    for i in range(B):
        feat[i, : feat_len[i]] = torch.randn(int(feat_len[i]), 80, generator=g) * 2 - 6
        tok[i, : tok_len[i]] = torch.randint(0, 4096, (int(tok_len[i]),), generator=g)
    return {'speech_token': tok, 'speech_token_len': tok_len, 'speech_feat': feat, 'speech_feat_len': feat_len,
            'embedding': torch.randn(B, 192, generator=g)}


def flow_model_leg(a, device, dtype, timed):
    """One optimiser step of the WHOLE flow model the reference trains in flow_only mode (flow_model.py:248-400:
    MaskedDiffWithXvec.forward(batch) -> loss.backward() -> clip + AdamW) with the reference's flow_lora target list
    (estimator attn1 q/k/v AND the Conformer encoder's linear_q/k/v, w_1, w_2; config.py:207-216). The estimator, the
    CFM and the optimiser run on the CUDA path; what sits in front of the estimator is reported separately so the
    cost of the path's callers is visible."""
    import random
    from cosyvoice_lora_finetune_framework_b200 import flow_model as FM, lora as LR, utils as U
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
    B, T = a.batch, a.frames
    U.set_all_random_seed(4321)
    m = FM.build_flow_model(None, 'cpu')
    LR.apply_lora_to_model(m, r=8, lora_alpha=16, lora_dropout=a.lora_dropout,
                           target_modules=['to_q', 'to_k', 'to_v', 'linear_q', 'linear_k', 'linear_v', 'w_1', 'w_2'])
    m = m.to(device).train()
    m.decoder.estimator.cvflow_dtype = dtype
    m.encoder_autocast = dtype            # the reference trains under Lightning '16-mixed' (config.py:76)
    upstream = [p for n, p in m.named_parameters() if p.requires_grad and not n.startswith('decoder.estimator.')]
    tr = FlowLoRATrainer(m.decoder, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0, extra_params=upstream)
    batch = flow_model_batch(B, T, 4321, a.min_len)
    random.seed(7)

    def step():
        out = m(batch, device)
        out['loss'].backward()
        tr.optimizer_step()
        return out['loss']

    for _ in range(3):
        step()
    n = 5
    ms = timed(step, n) / n
    loss = float(step())
    m.encoder_cuda_graphs = True          # encoder forward / backward replayed from CUDA graphs captured for this shape
    for _ in range(3):
        step()
    ms_g = timed(step, n) / n
    res = {"config": "MaskedDiffWithXvec.forward(batch) + backward + clip/AdamW, %d x %d frames, LoRA on estimator q/k/v and on "
                     "the 6-block Conformer encoder (%d upstream tensors), eager launches" % (B, T, len(upstream)),
           "ms_per_step": ms, "value": B * T / (ms / 1e3), "unit": UNIT, "loss": loss,
           "ms_per_step_encoder_cuda_graphs": ms_g,
           "path_inputs": getattr(m, "path_inputs_backend", "host PyTorch"),
           "encoder": "host PyTorch (eager, torch.autocast %s); SURVEY 8 f3 is not a kernel of this library" % str(dtype)}
    return res


def run_reference(a):
    """The reference's own algorithm for the path (oracle port) on all host cores, on the GPU arm's config. Each step
    is the FULL workload (32 x 400 frames, fp32, fwd + bwd + clip + AdamW, lora_dropout like the GPU arm) when that
    keeps the whole run within a few minutes; otherwise a smaller batch of the same shape, said so in the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    B_full = per_rank_batch(a, world)
    budget_s = 300.0
    # probe: one step at B = 2 (also loads the libraries), extrapolate linearly in B (CPU frames/s rises slowly with B)
    probe = oracle_step_fn(2, a.frames, "cpu", a.lora_dropout, a.min_len)
    probe()
    t2 = min(time_host(probe, 1))
    n_steps_total = max(1, a.steps) + max(0, a.warmup)
    B_s = B_full
    while B_s > 2 and (t2 * B_s / 2.0) * n_steps_total > budget_s:
        B_s //= 2
    step = oracle_step_fn(B_s, a.frames, "cpu", a.lora_dropout, a.min_len) if B_s != 2 else probe
    time_host(step, max(0, a.warmup))
    times = time_host(step, max(1, a.steps))
    total = sum(times)
    val = B_s * a.frames * len(times) / total
    cfg = workload_config(a, world)
    full = B_s == B_full
    sample = ("each step = %s: %d x %d frames (fp32, fwd + bwd + clip + AdamW, lora_dropout %g), %d host threads"
              % ("the full per-GPU workload" if full else "a bounded sample of the workload (batch cut to fit the time budget)",
                 B_s, a.frames, a.lora_dropout, cores))
    if not full:
        cfg["workload_sample"] = "reference arm ran batch %d x %d frames instead of %d x %d" % (B_s, a.frames, B_full, a.frames)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": len(times),
           "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
           "scaling": "strong" if a.global_batch else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": cfg, "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "sample_is_full_workload": full,
           "note": "reference algorithm restated in plain PyTorch fp32 (oracle/flow_oracle.py, pinned to the real "
                   "reference by tests/golden incl. this exact 32 x 400 batch) on all host cores; the reference itself is "
                   "pure Python/PyTorch and cannot travel to the GPU box"}
    print(json.dumps(out))


def _trace(msg):
    if os.environ.get("CVFLOW_BENCH_TRACE"):
        print("[bench rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


def run_cvflow(a):
    import torch.distributed as dist
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if os.environ.get("CVFLOW_BENCH_TRACE"):   # where is a hung rank? dump every thread's stack after a while
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("CVFLOW_BENCH_TRACE_AFTER", "75")), repeat=False, file=sys.stderr)
    from cosyvoice_lora_finetune_framework_b200 import _estimator as E
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _trace("process group ready")
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float16
    B, T, K, W = per_rank_batch(a, world), a.frames, a.steps, max(3, a.warmup)
    cfm, est, stats = build_model(device, dtype, a.lora_dropout)
    cfm.num_streams = max(1, a.streams)
    trainer = FlowLoRATrainer(cfm, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0)
    ne = trainer.ne
    batch, lens = make_batch(B, T, 99 + rank, device, a.min_len)
    torch.manual_seed(7 + rank)
    use_graph = not a.no_graph

    def mk_step(tr, b):
        if use_graph:   # the whole optimiser step (~1,300 launches) replayed as one CUDA graph
            return lambda: tr.train_step_graphed(b["x1"], b["mask"], b["mu"], b["spks"], b["cond"])
        return lambda: tr.train_step(b["x1"], b["mask"], b["mu"], b["spks"], b["cond"])

    step = mk_step(trainer, batch)
    eager_step = lambda: trainer.train_step(batch["x1"], batch["mask"], batch["mu"], batch["spks"], batch["cond"])

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn bracketed by barrier + synchronize on both sides, CUDA events, max over ranks; ms total."""
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    count = lambda: ne.launch_count() + sum(r.launch_count() for r in ne.replicas)
    _trace("model built")
    eager_step()
    torch.cuda.synchronize()
    _trace("first eager step done")
    l0 = count()
    eager_step()
    # + cfm_prep, loss (3), sumsq (2), optim_advance, adamw, merge (+ 6 batched torch copies of the un-folded LoRA images)
    per_step_launches = (count() - l0) + max(1, a.streams) * (1 + 3) + (2 + 1 + 1 + 1) + (6 if a.lora_dropout > 0 else 0)
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    _trace("warm-up (graph capture) done")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step, K)
    launches = K * per_step_launches
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * T * K / (ms / 1e3)
    _trace("timed region done: %.2f ms/step" % (ms / K))

    # ---- end to end through the public API with HOST buffers ---------------------------------------------
    # every step: its inputs are copied pinned host -> device (copy stream, overlapping the previous step's compute, the
    # way a pinned prefetching DataLoader feeds the reference trainer) and its loss is read back to the host
    host = {k: v.cpu().pin_memory() for k, v in batch.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    hargs = (host["x1"], host["mask"], host["mu"], host["spks"], host["cond"])

    def e2e_run(n):
        if not use_graph:
            for _ in range(n):
                dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
                float(trainer.train_step(dev["x1"], dev["mask"], dev["mu"], dev["spks"], dev["cond"]).item())
            return
        prev = None
        trainer.stage_inputs(*hargs)
        for i in range(n):
            if i + 1 < n:
                trainer.stage_inputs(*hargs)      # step i+1's host->device copy runs under step i
            h = trainer.train_step_staged()
            if prev is not None:
                prev.value()                      # step i-1's loss, read without stalling step i's launch
            prev = h
        prev.value()

    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(K), 1)
    _trace("e2e done")
    e2e_val = world * B * T * K / (ms_e2e / 1e3)

    # ---- per-kernel-class device time (CUDA events on the launching stream, eager steps) ---------------
    roofline, kernels = None, None
    L = E._lib()
    if rank == 0:
        L.cvflow_set_profile(ne.handle, 1)
    ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_a.record()
    for _ in range(2):          # every rank steps (the optimiser step holds the gradient allreduce); rank 0 records
        eager_step()
    ev_b.record()
    sync()
    ms_eager2 = ev_a.elapsed_time(ev_b)
    _trace("profiled eager steps done")
    peak_tf, peak_gbs, how = peaks()
    if rank == 0:
        n = 6
        msa, cnt, fl = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)()
        L.cvflow_profile_read(ne.handle, msa, cnt, fl, n)
        L.cvflow_set_profile(ne.handle, 0)
        spec = [("gemm_tc (tcgen05 implicit GEMM)", "tensor"), ("attn_fwd (tcgen05)", "tensor"),
                ("attn_bwd (tcgen05, dQ + dK/dV launches)", "tensor"), ("layernorm fwd+bwd (incl. fused LoRA-dropout forms)", "hbm"),
                ("lora_wgrad (tcgen05)", "tensor"), ("groupnorm+mish apply / bwd", "hbm")]
        kernels = {}
        for i, (name, bound) in enumerate(spec):
            if cnt[i] <= 0 or msa[i] <= 0:
                continue
            per_s = (fl[i] / 2) / (msa[i] / 2 * 1e-3)
            ach = per_s / 1e12 if bound == "tensor" else per_s / 1e9
            pk = peak_tf if bound == "tensor" else peak_gbs
            kernels[name] = {"ms_per_step": msa[i] / 2, "launches_per_step": int(cnt[i] // 2), "bound": bound,
                             "achieved": ach, "unit": "TFLOP/s" if bound == "tensor" else "GB/s", "peak": pk, "frac": ach / pk,
                             "avg_launch_us": 1e3 * msa[i] / max(1, cnt[i])}
        ach = (fl[0] / 2) / (msa[0] / 2 * 1e9)
        traffic, tsrc = ncu_gemm_traffic()
        roofline = {"kernel": "gemm_tc_kernel", "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "traffic": traffic, "traffic_source": tsrc, "peak_source": how,
                    "avg_launch_us": 1e3 * msa[0] / max(1, cnt[0]),
                    "share_of_eager_step": msa[0] / ms_eager2 if ms_eager2 > 0 else None,
                    "note": "algorithmic FLOPs = 2*M*N*K per launch with M = real (unpadded) rows, summed over the %d GEMM "
                            "launches of a step, / CUDA-event time of the same launches (eager step, launch gaps included); "
                            "HBM classes: algorithmic bytes (DESIGN.md section 4) / event time" % (cnt[0] // 2),
                    "whole_step": {"tflop_per_step": 228.54e6 * B * T / 1e12 if T == 400 else None,
                                   "frac_of_peak": (228.54e6 * B * T / 1e12) / (ms / K * 1e-3) / peak_tf if T == 400 else None}}

    extra = {}
    solo = rank == 0 and world == 1

    # ---- the folded-LoRA step (lora_dropout = 0: B A merged into the GEMM operand; the parity configuration) ----
    if solo and use_graph and not a.no_extra_legs and a.lora_dropout > 0:
        try:
            cfm3, _, _ = build_model(device, dtype, 0.0)
            tr3 = FlowLoRATrainer(cfm3, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0)
            step3 = mk_step(tr3, batch)
            for _ in range(W):
                step3()
            ms3 = timed(step3, K)
            extra["lora_dropout_0_leg"] = {"lora_dropout": 0.0, "ms_per_step": ms3 / K, "value": B * T * K / (ms3 / 1e3),
                                           "unit": UNIT, "note": "B A folded into the q/k/v operand (eval() / dropout 0)"}
            del tr3, cfm3, step3
            torch.cuda.empty_cache()
        except Exception as e:      # informational leg: never lose the headline line to it
            extra["lora_dropout_0_leg"] = {"error": repr(e)[:200]}
        _trace("folded leg done")

    # ---- configs[4]: long ragged utterances, 16 x 1500 frames per GPU ----
    if solo and use_graph and not a.no_extra_legs:
        try:
            B5, T5, K5 = 16, 1500, max(2, min(K, 4))
            b5, lens5 = make_batch(B5, T5, 199, device, 0.2)
            step5 = mk_step(trainer, b5)
            for _ in range(3):
                step5()
            ms5 = timed(step5, K5)
            extra["configs4_leg"] = {"config": "configs[4]: %d x %d frames per GPU, ragged lengths in (300, 1500]" % (B5, T5),
                                     "ms_per_step": ms5 / K5, "value": B5 * T5 * K5 / (ms5 / 1e3), "unit": UNIT,
                                     "valid_frames_per_step": int(lens5.sum()),
                                     "tflop_per_step": 377e6 * B5 * T5 / 1e12,
                                     "frac_of_peak": (377e6 * B5 * T5 / 1e12) / (ms5 / K5 * 1e-3) / peak_tf}
            trainer._graph = None
            del step5, b5
            torch.cuda.empty_cache()
        except Exception as e:
            extra["configs4_leg"] = {"error": repr(e)[:200]}
        _trace("configs[4] leg done")

    # ---- the reference algorithm in eager PyTorch on this GPU (bf16 autocast): "the existing Blackwell path" ----
    if solo and not a.no_extra_legs:
        try:
            fn = oracle_step_fn(B, T, device, a.lora_dropout, a.min_len, autocast=torch.bfloat16)
            for _ in range(2):
                fn()
            msg = timed(fn, 3)
            extra["eager_pytorch_gpu_leg"] = {"ms_per_step": msg / 3, "value": B * T * 3 / (msg / 1e3), "unit": UNIT,
                                              "note": "oracle port (plain PyTorch modules' math, cuBLAS/cuDNN/ATen kernels, bf16 "
                                                      "autocast, autograd, torch AdamW) on the same B200, same batch"}
            del fn
            torch.cuda.empty_cache()
        except Exception as e:
            extra["eager_pytorch_gpu_leg"] = {"error": repr(e)[:200]}
        _trace("eager GPU leg done")

    # ---- the whole flow model: token embedding -> Conformer encoder -> length regulator -> CFM ----
    if solo and not a.no_extra_legs:
        try:
            extra["flow_model_leg"] = flow_model_leg(a, device, dtype, timed)
        except Exception as e:
            extra["flow_model_leg"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
        _trace("flow model leg done")

    # ---- Euler-ODE inference (BASELINE configs[1]) -------------------------------------------------
    inference = None
    if solo and not a.no_inference:   # batch-1 by construction (replicas only): reported at N=1
        est.eval()
        Ti, P_, n_steps = 700, 200, 10
        g = torch.Generator().manual_seed(5)
        mu = torch.randn(1, 80, Ti, generator=g).to(device)
        spk = torch.randn(1, 80, generator=g).to(device)
        cond = torch.zeros(1, 80, Ti, device=device)
        cond[:, :, :P_] = torch.randn(1, 80, P_, generator=g).to(device)
        mask1 = torch.ones(1, 1, Ti, device=device)
        run = lambda: cfm(mu=mu.clone(), mask=mask1, n_timesteps=n_steps, spks=spk, cond=cond, prompt_len=P_)
        for _ in range(3):
            run()
        ms_inf = timed(run, 10) / 10
        audio_s = (Ti - P_) * 256 / 22050.0
        tflop = 2 * Ti * 119.04e6 * n_steps / 1e12
        inference = {"config": "configs[1]: 10 Euler steps + CFG, 500 target + 200 prompt frames, CUDA-graph replay",
                     "ms_per_solve": ms_inf, "rtf": (ms_inf / 1e3) / audio_s, "target_frames_per_s": (Ti - P_) / (ms_inf / 1e3),
                     "roofline": {"bound": "tensor", "achieved": tflop / (ms_inf * 1e-3), "peak": peak_tf, "unit": "TFLOP/s",
                                  "frac": tflop / (ms_inf * 1e-3) / peak_tf,
                                  "note": "whole solve: 1.67 TFLOP (BASELINE.md section 2) / solve time; M = 1,400 / 700 tokens per "
                                          "launch, launch-latency-bound"}}
        if not a.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            ce = oracle_euler_fn(Ti, P_, n_steps, "cpu")
            ce()
            tcpu = min(time_host(ce, 1))
            inference["cpu_baseline"] = {"value": tcpu / audio_s, "unit": "RTF", "seconds_per_solve": tcpu,
                                         "cores": torch.get_num_threads(), "kind": "port",
                                         "sample": "the full configs[1] solve (oracle port, fp32), best of 1 after 1 warm-up"}
        est.train()
        _trace("inference done")

    # ---- CPU baseline: the reference algorithm on this box's host cores, same shape ------------------
    cpu = None
    if solo and not a.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        warm = oracle_step_fn(2, T, "cpu", a.lora_dropout, a.min_len)
        warm()
        t2 = min(time_host(warm, 1))
        Bc = B
        while Bc > 2 and t2 * Bc / 2.0 > 30.0:       # keep the sample to ~10-30 s of CPU work
            Bc //= 2
        fn = oracle_step_fn(Bc, T, "cpu", a.lora_dropout, a.min_len) if Bc != 2 else warm
        best = min(time_host(fn, 1)) if Bc != 2 else t2
        cpu = {"value": Bc * T / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "one %d x %d-frame fp32 train step (fwd + bwd + clip + AdamW, lora_dropout %g) of the oracle port after a "
                         "2 x %d warm-up step; %.1f s" % (Bc, T, a.lora_dropout, T, best)}
        _trace("cpu baseline done")

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if a.global_batch else "weak",
               "vs_baseline": None, "dtype": a.dtype, "data": "synthetic", "config": workload_config(a, world), "clocks": clocks,
               "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                       "ms_per_step": ms_e2e / K,
                       "how": "pinned host inputs copied every step on a copy stream (step i+1's copy under step i's compute), "
                              "loss of every step read back through pinned memory one step late"},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels,
               "inference": inference, "valid_frames_per_step_rank0": int(lens.sum()),
               "parity": "bf16 operands / fp32 accumulation: loss <= 1e-2, LoRA grads within 1.5x the reference's own bf16-autocast "
                         "error at this exact batch (tests/test_train_gpu.py::test_benchmarked_shape_bf16, PARITY.md); fp16 "
                         "operands meet <= 1e-2 (test_benchmarked_shape_fp16)",
               "lora": {"replaced_layers": stats["replaced_layers"], "lora_params": stats["lora_params"],
                        "lora_dropout": a.lora_dropout}}
        out.update(extra)
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        # NCCL: a communicator whose collectives were captured into a CUDA graph must outlive that graph -- drop the
        # captured step first, and never let a teardown problem turn a finished benchmark into a hang
        sys.stdout.flush()
        dist.barrier()
        trainer._graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cvflow(args)
