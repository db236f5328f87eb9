#!/bin/bash
# round-2: the final default bench line at HEAD (all blocks)
set -u
o=gpurun_out
timeout 840 python bench.py > $o/r2w_bench_n1.json 2> $o/r2w_bench_n1.err
