"""One training step (configs[2]: 32 x 400, r=8) bracketed by cudaProfilerStart/Stop for ncu
(`ncu --profile-from-start off ...`). Also usable plain: prints the step time.
PROF_DROPOUT=0.05 runs the un-folded LoRA-dropout path, PROF_INPUT_GRADS=1 also asks for dL/dmu, dL/dspks, dL/dcond;
PROF_BLOCKS / PROF_MID shrink the estimator (same per-kernel sizes, fewer launches) for per-kernel ncu captures."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer  # noqa: E402

B = int(os.environ.get("PROF_B", "32"))
T = int(os.environ.get("PROF_T", "400"))
mode = os.environ.get("PROF_MODE", "train")


class A:
    batch, frames, dtype = B, T, "bf16"


dev = torch.device("cuda", 0)
drop = float(os.environ.get("PROF_DROPOUT", "0"))
if "PROF_BLOCKS" in os.environ:
    from cosyvoice_lora_finetune_framework_b200 import lora, modules
    from cosyvoice_lora_finetune_framework_b200.flow_model import ConditionalCFM
    est = modules.ConditionalDecoder(in_channels=320, out_channels=80, channels=(256, 256), dropout=0.0, attention_head_dim=64,
                                     n_blocks=int(os.environ["PROF_BLOCKS"]), num_mid_blocks=int(os.environ.get("PROF_MID", "1")),
                                     num_heads=8, act_fn='gelu')
    lora.apply_lora_to_model(est, r=8, lora_alpha=16, lora_dropout=drop, target_modules=['to_q', 'to_k', 'to_v'])
    est = est.to(dev).train()
    est.cvflow_dtype = torch.bfloat16
    cfm = ConditionalCFM(in_channels=80, n_spks=1, spk_emb_dim=80, estimator=est)
else:
    cfm, est, _ = bench.build_model(dev, torch.bfloat16, drop)
if mode == "train":
    tr = FlowLoRATrainer(cfm)
    batch, _ = bench.make_batch(B, T, 99, dev)
    if os.environ.get("PROF_INPUT_GRADS"):
        for k in ("mu", "spks", "cond"):
            batch[k].requires_grad_(True)
    step = lambda: tr.train_step(batch["x1"], batch["mask"], batch["mu"], batch["spks"], batch["cond"])
else:
    est.eval()
    cfm.use_cuda_graph = False
    mu = torch.randn(1, 80, T, device=dev)
    step = lambda: cfm(mu=mu.clone(), mask=torch.ones(1, 1, T, device=dev), n_timesteps=2, spks=torch.randn(1, 80, device=dev),
                       cond=torch.zeros(1, 80, T, device=dev))
for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
t0 = time.perf_counter()
step()
torch.cuda.synchronize()
t1 = time.perf_counter()
torch.cuda.profiler.stop()
print("step ms", (t1 - t0) * 1e3)
