#!/bin/bash
# pose solver: concurrently solved pairs (CTA groups) vs L2 residency of the per-pair inputs (13.76 MB each, re-read every evaluation)
set -u
O=gpurun_out/${1:-s15}
mkdir -p $O
for g in 4 6 8 12 16; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --latency-frames 0 --pose-groups $g > $O/bench_groups$g.json 2> $O/bench_groups$g.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_groups$g.json")); s=d["stages"]["pose_solve"]
    print("groups $g: value %.1f ms/step %.1f pose_solve %.2f ms/step (%.0f us/launch)" % (d["value"], d["ms_per_step"], s["total_ms"]/d["steps"], s["avg_us"]))
except Exception as e: print("groups $g failed", e)
PY
done
