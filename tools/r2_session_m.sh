#!/bin/bash
set -u
O=gpurun_out/r2_m
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_update_tc.py tests/test_gpu_e2e.py -m gpu -q --timeout 600 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.1f e2e %.1f ms/step %.1f clocks %s parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["pose_parity"]["max_rel_translation"]))
for k,v in d["stages"].items(): print("%-18s launches %5d total_ms %9.2f avg_us %9.1f min %9.1f max %9.1f" % (k, v["launches"], v["total_ms"], v["avg_us"], v["min_us"], v["max_us"]))
PY
