#!/bin/bash
# round-2: ncu launch list (DRAM bytes) and hardware-counter sections of the benched replay launch on the FINAL default workload
set -u
o=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
   --log-file $o/r2n_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-full-job --no-config5 > $o/r2n_ncu_l.log 2>&1
timeout 420 ncu --replay-mode application --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy \
   --section LaunchStats --clock-control none -k regex:k_anneal_replay -c 1 \
   -o $o/r2n_replay_sections -f python tools/profile_run.py --workload c3 --reads 75776 --sweeps 50 > $o/r2n_ncu_replay.log 2>&1
