import os, sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from cosyvoice_lora_finetune_framework_b200.trainer import FlowLoRATrainer
class A: batch, frames, dtype = 32, 400, "bf16"
dev = torch.device("cuda", 0)
cfm, est, _ = bench.build_model(A, dev, torch.bfloat16)
tr = FlowLoRATrainer(cfm)
batch, _ = bench.make_batch(32, 400, 99, dev)
step = lambda: tr.train_step(batch["x1"], batch["mask"], batch["mu"], batch["spks"], batch["cond"])
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue ms/step %.2f ; total ms/step %.2f" % ((t1 - t0) / 5 * 1e3, (t2 - t0) / 5 * 1e3))
