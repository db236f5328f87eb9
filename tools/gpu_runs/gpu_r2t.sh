#!/bin/bash
# round-2: two INDEPENDENT single-GPU processes (no torch.distributed, no NCCL), one per GPU, each annealing one half of the stated
# job: does the 45 % slow-down of one rank at N = 2 come from the box (two busy GPUs) or from the process group?
set -u
o=gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-config5 --job-share 0/2 > $o/r2t_gpu0.json 2> $o/r2t_gpu0.err &
P0=$!
CUDA_VISIBLE_DEVICES=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-config5 --job-share 1/2 > $o/r2t_gpu1.json 2> $o/r2t_gpu1.err &
P1=$!
wait $P0 $P1
