#!/bin/bash
# round-2: rank 1's share of the N = 2 strong job (reads 50 000 .. 99 999, its seeds and initial states) on ONE GPU: is the 120 s of
# the slower rank a property of its data or of the second GPU / process?  (also: the GPU suite after the CTA-width rule change)
set -u
o=gpurun_out
timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-config5 --job-share 1/2 > $o/r2s_bench.json 2> $o/r2s_bench.err
timeout 600 python -m pytest tests -q -m gpu -x > $o/r2s_tests.log 2>&1
