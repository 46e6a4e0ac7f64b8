#!/bin/bash
# Round 2, GPU session C: fp16 hi + bf16 lo planes (mixed-format MMAs), heads on the conv kernels, lookup v2: whole suite + bench
set -u
O=gpurun_out/r2_c
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^heads |config1|bench64|sharded|only3d|passed|failed|FAILED|corr fp16x3|lookup planes|Error" $O/pytest_gpu.log | head -70
timeout 900 python bench.py --no-cpu-baseline --config5-frames 0 --latency-pairs 40 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json"))
    print("value %.1f e2e %.1f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
    print("parity", d["pose_parity"]); print("gpu_reference", d.get("gpu_reference",{}).get("value"))
    for k,v in d["stages"].items(): print("%-18s launches %5d total_ms %9.2f avg_us %9.1f" % (k, v["launches"], v["total_ms"], v["avg_us"]))
    for k,v in d["kernels"].items(): print("%-26s %s achieved %8.1f frac %.3f" % (k, v["unit"], v["achieved"], v["frac"]))
except Exception as e:
    print("bench parse failed", e)
PY
