#!/bin/bash
set -u
O=gpurun_out/r2_cv
mkdir -p $O
timeout 600 python tools/corr_variance_probe.py 5 > $O/corr_variance.txt 2>&1; cat $O/corr_variance.txt | tail -14
