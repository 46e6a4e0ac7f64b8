#!/bin/bash
set -u
O=gpurun_out/r2_f
mkdir -p $O
for i in 1 2; do
timeout 600 python -m pytest tests/test_gpu_update_tc.py -m gpu -q -s --timeout 600 -k "update_operator" 2>&1 | grep -E "update operator|fused vs plain|passed|failed" | tee -a $O/rerun.txt
done
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_update_tc.py -m gpu -q -s --timeout 600 -k "update_operator" 2>&1 | grep -E "update operator|fused vs plain|passed|failed" | tee -a $O/rerun.txt
RPE_CONV_PAIR=0 timeout 600 python -m pytest tests/test_gpu_update_tc.py -m gpu -q -s --timeout 600 -k "update_operator" 2>&1 | grep -E "update operator|fused vs plain|passed|failed" | tee -a $O/rerun.txt
