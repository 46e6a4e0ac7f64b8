#!/bin/bash
set -u
O=gpurun_out/r2_s
mkdir -p $O
for m in 1 2 3; do
  echo "== RPE_CONV_BGROUP=$m" >> $O/conv_probe_bgroup.txt
  RPE_CONV_BGROUP=$m timeout 300 python tools/conv_probe.py --n=32 >> $O/conv_probe_bgroup.txt 2>&1
done
cat $O/conv_probe_bgroup.txt
