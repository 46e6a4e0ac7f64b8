#!/bin/bash
# compute-sanitizer: memcheck over the end-to-end GPU tests, synccheck + racecheck over the convolution / correlation unit tests
set -u
O=gpurun_out/r2_san
mkdir -p $O
timeout 1800 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_configs.py tests/test_variants.py tests/test_file_dataset.py -m gpu -q -x --timeout 1700 > $O/memcheck_e2e.log 2>&1; echo "memcheck e2e rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" $O/memcheck_e2e.log | head -6
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_update_tc.py tests/test_gpu_corr.py -m gpu -q -x --timeout 800 > $O/synccheck_conv.log 2>&1; echo "synccheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" $O/synccheck_conv.log | head -6
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_update_tc.py -m gpu -q -x --timeout 1100 -k "conv_fp16x3_vs_torch or residual" > $O/racecheck_conv.log 2>&1; echo "racecheck rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/racecheck_conv.log | head -6
