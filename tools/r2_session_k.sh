#!/bin/bash
set -u
O=gpurun_out/r2_k
mkdir -p $O
timeout 600 python tools/pose_probe.py 32 2>&1 | grep -v Warn | tee $O/pose_probe.txt
timeout 600 python tools/pose_probe.py 1 2>&1 | grep -v Warn | tee -a $O/pose_probe.txt
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_corr.py -m gpu -q --timeout 600 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 40 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json"))
    print("value %.1f e2e %.1f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
    print("parity", d["pose_parity"]["max_rel_translation"], "latency", d["latency"]["p50_ms_per_pair"], d["latency"]["eager"]["p50_ms_per_pair"], "clocks", d["clocks"])
    for k,v in d["stages"].items(): print("%-18s launches %5d total_ms %9.2f avg_us %9.1f" % (k, v["launches"], v["total_ms"], v["avg_us"]))
except Exception as e:
    print("bench parse failed", e)
PY
