#!/bin/bash
set -u
O=gpurun_out/r2_w
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error|residual from planes" $O/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 40 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.1f e2e %.1f ms/step %.1f clocks %s parity %s latency %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["pose_parity"], d["latency"]["p50_ms_per_pair"]))
for k,v in d["stages"].items(): print("%-18s launches %5d total_ms %9.2f avg_us %9.1f min %9.1f max %9.1f" % (k, v["launches"], v["total_ms"], v["avg_us"], v["min_us"], v["max_us"]))
PY

