#!/bin/bash
set -u
O=gpurun_out/r2_d
mkdir -p $O
timeout 600 python tools/pose_probe.py 32 2>&1 | grep -v Warn | tee $O/pose_probe.txt
timeout 600 python tools/pose_probe.py 8 2>&1 | grep -v Warn | tee -a $O/pose_probe.txt
timeout 600 python tools/pose_probe.py 1 2>&1 | grep -v Warn | tee -a $O/pose_probe.txt
timeout 900 python -m pytest tests/test_gpu_heads.py tests/test_gpu_stages.py -m gpu -q -s --timeout 600 2>&1 | grep -E "^heads|passed|failed|Error" | tee $O/pytest_heads.log
