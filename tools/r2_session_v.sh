#!/bin/bash
set -u
O=gpurun_out/r2_v
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_update_tc.py tests/test_gpu_heads.py -m gpu -q -x --timeout 600 > $O/pytest_conv.log 2>&1; echo "pytest conv rc=$?"; tail -4 $O/pytest_conv.log
for m in 0 96 128; do
  echo "== RPE_CONV_MERGE=$m" >> $O/conv_probe_merge.txt
  RPE_CONV_MERGE=$m timeout 300 python tools/conv_probe.py --n=32 >> $O/conv_probe_merge.txt 2>&1
done
cat $O/conv_probe_merge.txt
