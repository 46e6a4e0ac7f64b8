#!/bin/bash
set -u
O=gpurun_out/r2_final2
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0"
python - <<PY
PY
timeout 900 $B > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.1f e2e %.1f ms/step %.1f clocks %s parity %.3e" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["pose_parity"]["max_rel_translation"]))
for k in ("conv_tc","conv_tc_enc","corr_build","pose_solve"): v=d["stages"][k]; print("%-18s total_ms %9.2f avg_us %9.1f" % (k, v["total_ms"], v["avg_us"]))
PY
