"""Where one optimiser step of the WHOLE flow model (bench.py flow_model_leg: MaskedDiffWithXvec.forward(batch) ->
backward -> clip + AdamW, LoRA on the estimator and on the Conformer encoder) spends its time.

Two views: (a) wall clock per phase with a device synchronize between phases (host enqueue + device time of the phase,
no overlap between phases), (b) the un-instrumented step (CUDA events). Run on a B200:
    python profiles/flow_model_breakdown.py > profiles/r02_flow_model_breakdown.txt"""
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from cosyvoice_lora_finetune_framework_b200 import flow_model as FM, lora as LR, utils as U
    from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
    dev = torch.device("cuda:0")
    B, T = 32, 400
    U.set_all_random_seed(4321)
    m = FM.build_flow_model(None, 'cpu')
    LR.apply_lora_to_model(m, r=8, lora_alpha=16, lora_dropout=0.05,
                           target_modules=['to_q', 'to_k', 'to_v', 'linear_q', 'linear_k', 'linear_v', 'w_1', 'w_2'])
    m = m.to(dev).train()
    m.decoder.estimator.cvflow_dtype = torch.bfloat16
    m.encoder_autocast = torch.bfloat16 if os.environ.get('FMB_AUTOCAST', '1') == '1' else None
    m.encoder_cuda_graphs = os.environ.get('FMB_ENC_GRAPHS', '0') == '1'
    upstream = [p for n, p in m.named_parameters() if p.requires_grad and not n.startswith('decoder.estimator.')]
    tr = FlowLoRATrainer(m.decoder, lr=1e-4, weight_decay=0.01, max_grad_norm=1.0, extra_params=upstream)
    batch = bench.flow_model_batch(B, T, 4321)
    random.seed(7)
    phases = {}
    cfm = m.decoder
    real_compute_loss = cfm.compute_loss
    marks = {}

    def wall(tag, t0):
        torch.cuda.synchronize()
        phases.setdefault(tag, []).append((time.perf_counter() - t0) * 1e3)

    def compute_loss(*a, **k):
        wall("1 path inputs: embedding, encoder, regulator, speaker, prompt loop (forward)", marks["t"])
        t0 = time.perf_counter()
        out = real_compute_loss(*a, **k)
        wall("2 compute_loss forward (CFM prep + estimator + loss, CUDA path)", t0)
        return out

    def step(instrument):
        cfm.compute_loss = compute_loss if instrument else real_compute_loss
        marks["t"] = time.perf_counter()
        out = m(batch, dev)
        t0 = time.perf_counter()
        out['loss'].backward()
        if instrument:
            wall("3 backward (estimator CUDA path + autograd through regulator / encoder)", t0)
        t0 = time.perf_counter()
        tr.optimizer_step()
        if instrument:
            wall("4 clip + AdamW over both buckets + LoRA refresh", t0)

    for _ in range(3):
        step(False)
    torch.cuda.synchronize()
    for _ in range(5):
        step(True)
    print("phase (wall clock incl. synchronize, median of 5, ms)")
    tot = 0.0
    for k in sorted(phases):
        v = sorted(phases[k])[len(phases[k]) // 2]
        tot += v
        print("  %-90s %8.2f" % (k, v))
    print("  %-90s %8.2f" % ("sum", tot))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 5
    for _ in range(n):
        step(False)
    e1.record()
    torch.cuda.synchronize()
    print("un-instrumented step (CUDA events): %.2f ms" % (e0.elapsed_time(e1) / n))
    print("path inputs backend:", getattr(m, "path_inputs_backend", "host PyTorch"), "| encoder autocast:", m.encoder_autocast, "| encoder CUDA graphs:", m.encoder_cuda_graphs)
    # host-only cost of the estimator part: the same step with prepared inputs through the trainer (eager, then graph)
    b, _ = bench.make_batch(B, T, 99, dev)
    tr2 = FlowLoRATrainer(cfm, lr=1e-4)
    for name, fn in (("estimator-only step, eager launches", lambda: tr2.train_step(b["x1"], b["mask"], b["mu"], b["spks"], b["cond"])),
                     ("estimator-only step, whole-step CUDA graph", lambda: tr2.train_step_graphed(b["x1"], b["mask"], b["mu"], b["spks"], b["cond"]))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("%s: %.2f ms" % (name, e0.elapsed_time(e1) / n))


if __name__ == "__main__":
    main()
