#!/bin/bash
# compute-sanitizer memcheck over smoke() and the small-shape GPU tests (kernel memory errors would show as ERROR SUMMARY != 0)
set -u
O=gpurun_out/r2_san
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py --smoke > $O/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
grep -E "ERROR SUMMARY|Invalid|smoke ok" $O/memcheck_smoke.log | head -10
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_update_tc.py tests/test_gpu_corr.py tests/test_gpu_stages.py tests/test_gpu_heads.py tests/test_gpu_preproc.py tests/test_resize.py -m gpu -q -x --timeout 1200 > $O/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" $O/memcheck_tests.log | head -10
