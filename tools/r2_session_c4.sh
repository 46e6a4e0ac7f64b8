#!/bin/bash
set -u
O=gpurun_out/r2_c4
mkdir -p $O
timeout 600 python tools/config4_probe.py 16 8 > $O/config4.txt 2>&1; tail -5 $O/config4.txt
