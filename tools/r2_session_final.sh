#!/bin/bash
# Round-end rehearsal of what the driver runs: smoke(), the default bench line, the reference arm -- with the wall clock of each.
set -u
O=gpurun_out/r2_final
mkdir -p $O
t0=$(date +%s); timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s); timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
t0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$? ($(( $(date +%s) - t0 )) s)"
tail -1 $O/smoke.log; cut -c1-400 $O/bench_reference.json; echo
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("value %.1f e2e %.1f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
print("latency", d["latency"]["p50_ms_per_pair"], d["latency"]["eager"]["p50_ms_per_pair"], "cpu_baseline", d.get("cpu_baseline"))
print("gpu_reference", d.get("gpu_reference"))
print("config5", d.get("config5"))
print("roofline", d["roofline"])
print("clocks", d["clocks"], "parity", d["pose_parity"])
PY
