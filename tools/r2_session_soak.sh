#!/bin/bash
# soak: 300 timed steps (19200 pairs) + a 5000-frame sequence end to end; parity of the last step and failed-pair counts
set -u
O=gpurun_out/r2_soak
mkdir -p $O
timeout 900 python bench.py --steps 300 --warmup 3 --no-cpu-baseline --no-gpu-reference --latency-pairs 0 --config5-frames 5000 > $O/bench_soak.json 2> $O/bench_soak.err; echo "soak rc=$?"; tail -2 $O/bench_soak.err
python - <<PY
import json
d=json.load(open("$O/bench_soak.json"))
print("steps %d value %.1f e2e %.1f ms/step %.2f clocks %s failed %s parity %s" % (d["steps"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"], d["e2e"]["failed_pairs"], d["pose_parity"]))
c=d["config5"]; print("config5 frames %d pairs/s %.1f seconds %.2f failed %d ate %.3f mm" % (c["frames"], c["pairs_per_s"], c["seconds"], c["failed_pairs"], c["accuracy"]["ate_rmse_mm"]))
PY
