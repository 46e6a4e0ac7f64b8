#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -q --timeout 200 2>&1 | tail -12
