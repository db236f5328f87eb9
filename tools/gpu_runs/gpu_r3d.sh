#!/bin/bash
set -u
timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/r3d_tests.log 2>&1
