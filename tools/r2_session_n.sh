#!/bin/bash
set -u
O=gpurun_out/r2_n
mkdir -p $O
echo "== default (both operands streaming, streaming stores)"; timeout 300 python tools/corr_probe.py 64 2>&1 | grep -E "planes in|MHz" | tee $O/a.txt
echo "== plain stores"; RPE_CORR_PLAIN_STORES=1 timeout 300 python tools/corr_probe.py 64 2>&1 | grep -E "planes in|MHz" | tee $O/b.txt
echo "== resident query tile"; RPE_CORR_RESIDENT_A=1 timeout 300 python tools/corr_probe.py 64 2>&1 | grep -E "planes in|MHz" | tee $O/c.txt
