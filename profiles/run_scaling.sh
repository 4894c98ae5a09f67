#!/bin/bash
# Weak / strong scaling and the configs[4] long-utterance sweep on one 8-GPU box (gpurun --gpus 8).
# Every line of gpurun_out/r02_scaling.jsonl is one bench.py JSON line (device time, max over ranks).
out=gpurun_out/r02_scaling.jsonl
: > $out
run() {  # nproc, extra args...
  n=$1; shift
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 8 --warmup 3 --no-inference --no-cpu-baseline --no-extra-legs "$@" >> $out 2>> gpurun_out/r02_scaling.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 200)) \
      bench.py --gpus $n --steps 8 --warmup 3 "$@" >> $out 2>> gpurun_out/r02_scaling.err
  fi
}
run 1                                  # weak, N = 1 (32 x 400 per GPU)
run 8                                  # weak, N = 8
run 4 --global-batch 32                # strong: 8 utterances per GPU
run 8 --global-batch 32                # strong: 4 utterances per GPU
run 8 --frames 1500 --min-len 0.2 --batch 1     # configs[4]: global batch 8
run 8 --frames 1500 --min-len 0.2 --batch 4     # global batch 32
run 8 --frames 1500 --min-len 0.2 --batch 16    # global batch 128
