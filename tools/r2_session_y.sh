#!/bin/bash
set -u
O=gpurun_out/r2_y
mkdir -p $O
for m in 0 1 0 1; do
  echo "== RPE_CONV_SIDE_PREFETCH=$m" >> $O/conv_probe_side_prefetch.txt
  RPE_CONV_SIDE_PREFETCH=$m timeout 300 python tools/conv_probe.py --n=64 --only=zr --only=q1 --only=q2 >> $O/conv_probe_side_prefetch.txt 2>&1
done
cat $O/conv_probe_side_prefetch.txt
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
for m in 0 1; do
RPE_CONV_SIDE_PREFETCH=$m timeout 900 python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0 > $O/bench_sp$m.json 2> $O/bench_sp$m.err; echo "bench sp=$m rc=$?"; tail -2 $O/bench_sp$m.err
python - <<PY
import json
d=json.load(open("$O/bench_sp$m.json"))
print("side_prefetch=$m value %.1f e2e %.1f ms/step %.1f clocks %s parity %.3e" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["pose_parity"]["max_rel_translation"]))
for k in ("conv_tc","conv_tc_enc","corr_build"): v=d["stages"][k]; print("%-18s total_ms %9.2f avg_us %9.1f" % (k, v["total_ms"], v["avg_us"]))
PY
done
