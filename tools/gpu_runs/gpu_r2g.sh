#!/bin/bash
# round-2 check: GPU tests, one short bench (parity diagnostics), ncu launch list + section capture of the benched launch
set -u
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $o/r2g_tests.log 2>&1
timeout 900 python bench.py --no-full-job > $o/r2g_bench.json 2> $o/r2g_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
   --log-file $o/r2g_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-full-job > $o/r2g_ncu_l.log 2>&1
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:k_anneal_replay -c 1 -o $o/r2g_replay_full -f \
   python tools/profile_run.py --reads 75776 --sweeps 50 > $o/r2g_ncu_full.log 2>&1
ls -la $o/*.ncu-rep >> $o/r2g_ncu_full.log 2>&1
