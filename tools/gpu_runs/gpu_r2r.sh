#!/bin/bash
# round-2: the strong leg at N = 2 again, with per-rank clocks / throttle reasons / per-rank durations in the `strong` block
set -u
o=gpurun_out
(while true; do nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown --format=csv,noheader >> $o/r2r_clocks.log; sleep 2; done) &
SMI=$!
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
   bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --no-config5 --no-cpu-baseline > $o/r2r_bench_n2.json 2> $o/r2r_bench_n2.err
kill $SMI
