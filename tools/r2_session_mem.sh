#!/bin/bash
set -u
O=gpurun_out/r2_mem
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "gpu memory|passed|failed" $O/pytest_gpu.log
