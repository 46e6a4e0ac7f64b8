#!/bin/bash
set -u
O=gpurun_out/r2_san
mkdir -p $O
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_update_tc.py -m gpu -q -x --timeout 1100 -k "conv_fp16x3_vs_torch or residual or instnorm_partial" > $O/racecheck_conv2.log 2>&1; echo "racecheck rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/racecheck_conv2.log | head -6
grep -A3 "Race reported" $O/racecheck_conv2.log | head -12
timeout 600 python -m pytest tests/test_gpu_update_tc.py tests/test_gpu_e2e.py -m gpu -q --timeout 500 2>&1 | tail -2
