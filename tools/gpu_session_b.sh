#!/bin/bash
# Short GPU call: parity tests, one bench line, conv probe.
set -u
O=gpurun_out/${1:-s6}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rs -s -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
timeout 300 python tools/conv_probe.py --n=22 > $O/probe.txt 2>&1
tail -5 $O/pytest_gpu.log; cut -c1-400 $O/bench.json; cat $O/probe.txt
