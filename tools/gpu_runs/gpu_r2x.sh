#!/bin/bash
# round-2: configs 2 and 4 on the final library (CTA-width rule changed since the last probes)
set -u
o=gpurun_out
timeout 200 python tools/probe_c2.py > $o/r2x_c2.log 2>&1
timeout 300 python tools/probe_c4.py > $o/r2x_c4.log 2>&1
