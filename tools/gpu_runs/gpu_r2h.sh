#!/bin/bash
# round-2: new components (dense tensor-core kernel, SNN builder, device recursion) + DMMA peak + config-5 probe
set -u
o=gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_dmma tools/ubench_dmma.cu && timeout 120 /tmp/ubench_dmma > $o/r2h_dmma.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_snn.py tests/test_gpu_recursion.py -q -m gpu > $o/r2h_new_tests.log 2>&1
timeout 600 python tools/probe_c5.py > $o/r2h_c5.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_dense.py --deselect tests/test_gpu_snn.py --deselect tests/test_gpu_recursion.py > $o/r2h_tests.log 2>&1
