#!/bin/bash
# Round-end rehearsal of what the driver runs: smoke(), the default bench line (with cpu_baseline and latency), the reference arm.
set -u
O=gpurun_out/${1:-final}
mkdir -p $O
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
tail -2 $O/smoke.log; cut -c1-300 $O/bench_default.json; echo; cut -c1-400 $O/bench_reference.json
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("value %.1f e2e %.1f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
print("latency", d["latency"]["p50_ms_per_pair"], "cpu_baseline", d.get("cpu_baseline"))
print("roofline", d["roofline"]["frac"], d["roofline"]["executed_frac"], "clocks", d["clocks"])
PY
