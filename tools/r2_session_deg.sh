#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 300 -k "fully_masked or edge_cases" 2>&1 | tail -12
