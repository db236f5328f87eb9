#!/bin/bash
# round-2 final (1 GPU): default bench line with the SA-friendly default size penalty (B = A / n), CQM-related GPU tests
set -u
o=gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_fullsize.py tests/test_gpu_dense.py tests/test_gpu_sampler.py -q -m gpu > $o/r2m_tests.log 2>&1
timeout 1500 python bench.py > $o/r2m_bench_n1.json 2> $o/r2m_bench_n1.err
