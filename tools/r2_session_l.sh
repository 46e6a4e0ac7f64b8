#!/bin/bash
set -u
O=gpurun_out/r2_l
mkdir -p $O
for v in cs plain cs plain; do
  if [ $v = plain ]; then export RPE_CORR_PLAIN_STORES=1; else unset RPE_CORR_PLAIN_STORES; fi
  timeout 900 python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0 > $O/bench_$v.json 2> $O/bench_$v.err
  python - <<PY
import json
d=json.load(open("$O/bench_$v.json"))
s=d["stages"]
print("$v value %.1f ms/step %.1f corr_build %.0f us lookup %.0f us pose %.0f us conv_tc %.1f clocks %s" % (d["value"], d["ms_per_step"], s["corr_build"]["avg_us"], s["corr_lookup"]["avg_us"], s["pose_solve"]["avg_us"], s["conv_tc"]["avg_us"], d["clocks"]["sm_mhz"]))
PY
done
