#!/bin/bash
set -u
timeout 240 python tools/probe_recursion.py 8192 > gpurun_out/r2z_recursion.log 2>&1
