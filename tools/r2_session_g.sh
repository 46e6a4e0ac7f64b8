#!/bin/bash
set -u
O=gpurun_out/r2_g
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^heads |config1|bench64|sharded|only3d|passed|failed|FAILED|Error" $O/pytest_gpu.log | head -40
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json"))
    print("value %.1f e2e %.1f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
    print("parity", d["pose_parity"]); print("latency", d["latency"]["p50_ms_per_pair"], d["latency"]["eager"]["p50_ms_per_pair"])
    print("config5", d["config5"]); print("gpu_reference", d.get("gpu_reference")); print("cpu_baseline", d.get("cpu_baseline"))
    print("roofline", d["roofline"]["frac"], d["roofline"]["executed_frac"], "clocks", d["clocks"])
    tot=0
    for k,v in d["stages"].items():
        print("%-18s launches %5d total_ms %9.2f avg_us %9.1f" % (k, v["launches"], v["total_ms"], v["avg_us"]))
        if k not in ("heads",): tot+=v["total_ms"]
    print("sum of stage timers per step %.1f of %.1f ms" % (tot/d["steps"], d["ms_per_step"]))
    for k,v in d["kernels"].items(): print("%-26s %s achieved %8.1f frac %.3f" % (k, v["unit"], v["achieved"], v["frac"]))
except Exception as e:
    print("bench parse failed", e)
PY
