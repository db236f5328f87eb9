#!/bin/bash
# round-2: the stated job's per-rank share at N = 2 (50 000 reads x 1000 sweeps) on ONE GPU -- separates the few-reads effect from
# anything specific to two busy GPUs in one box -- with clocks sampled during the job
set -u
o=gpurun_out
(while true; do nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader >> $o/r2q_clocks.log; sleep 2; done) &
SMI=$!
timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-config5 --job-reads 50000 > $o/r2q_bench.json 2> $o/r2q_bench.err
kill $SMI
