#!/bin/bash
# round-2: warps per CTA of the replay kernel at the read counts of the strong-scaled job (100 000 reads over 2 / 4 / 8 GPUs),
# plus the new CQM feasibility test
set -u
o=gpurun_out
timeout 300 python -m pytest tests/test_gpu_sampler.py -q -m gpu -k "cqm or dqm" > $o/r2p_tests.log 2>&1
: > $o/r2p_warps.log
for reads in 50000 25000 12500; do
  echo "== reads $reads" >> $o/r2p_warps.log
  timeout 400 python tools/probe_replay.py --reads $reads --permille 20 --warps 8,4,2,1 2>&1 | grep -E "sweeps|Error|error" >> $o/r2p_warps.log
done
