#!/bin/bash
set -u
timeout 300 python tools/probe_snn.py > gpurun_out/r2y_snn.log 2>&1
