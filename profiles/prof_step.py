"""One training step (configs[2]: 32 x 400, r=8) bracketed by cudaProfilerStart/Stop for ncu
(`ncu --profile-from-start off ...`). Also usable plain: prints the step time."""
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
cfm, est, _ = bench.build_model(A, dev, torch.bfloat16)
if mode == "train":
    tr = FlowLoRATrainer(cfm)
    batch, _ = bench.make_batch(B, T, 99, dev)
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
