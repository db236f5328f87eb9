#!/bin/bash
set -u
o=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_snn.py tests/test_gpu_recursion.py tests/test_gpu_sampler.py -q -m gpu > $o/r2i_new_tests.log 2>&1
timeout 600 python tools/probe_c5.py --ref-reads 0 > $o/r2i_c5.log 2>&1
