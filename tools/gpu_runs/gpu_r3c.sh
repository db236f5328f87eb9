#!/bin/bash
set -u
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_replay.py -q -m gpu -k "one_shot or interrupt" > gpurun_out/r3c_tests.log 2>&1
