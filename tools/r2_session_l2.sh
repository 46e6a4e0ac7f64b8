#!/bin/bash
set -u
O=gpurun_out/r2_l2
mkdir -p $O
timeout 600 python tools/l2_fetch_probe.py > $O/l2_fetch.txt 2>&1; tail -12 $O/l2_fetch.txt
