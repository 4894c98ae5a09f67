#!/bin/bash
# Weak scaling N = 1 / 8 and configs[4] (16 x 1500 per GPU) with the library at the end of the round (gpurun --gpus 8).
out=gpurun_out/r02_scaling_final.jsonl
: > $out
python bench.py --gpus 1 --steps 8 --warmup 3 --no-inference --no-cpu-baseline --no-extra-legs >> $out 2>> gpurun_out/r02_scaling_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --steps 8 --warmup 3 >> $out 2>> gpurun_out/r02_scaling_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 8 --steps 8 --warmup 3 --frames 1500 --min-len 0.2 --batch 16 >> $out 2>> gpurun_out/r02_scaling_final.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_scaling_final.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print(d["n_gpus"], d["config"]["frames"], d["config"]["global_batch"], round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]))
PY
