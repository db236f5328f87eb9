#!/bin/bash
# round-2: N = 2 strong leg with NCCL's peer-to-peer / NVLS paths disabled -- does the one slow rank come from peer mappings?
set -u
o=gpurun_out
NCCL_NVLS_ENABLE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
   bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --no-config5 --no-cpu-baseline > $o/r2v_bench_n2.json 2> $o/r2v_bench_n2.err
