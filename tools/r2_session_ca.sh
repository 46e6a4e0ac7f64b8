#!/bin/bash
set -u
O=gpurun_out/r2_ca
mkdir -p $O
timeout 600 python tools/corr_align_probe.py > $O/corr_align.txt 2>&1; cat $O/corr_align.txt | tail -24
