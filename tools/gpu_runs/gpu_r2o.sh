#!/bin/bash
# round-2 final (2 GPUs): the bench as the driver launches it for N = 2 (weak `value`, `strong` block, e2e legs, config5)
set -u
o=gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 2 --steps 2 --warmup 3 > $o/r2o_bench_n2.json 2> $o/r2o_bench_n2.err
tail -c 1500 $o/r2o_bench_n2.err
