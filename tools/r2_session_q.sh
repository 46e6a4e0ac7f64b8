#!/bin/bash
set -u
O=gpurun_out/r2_q
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
for g in "" "--graphs"; do
timeout 600 python bench.py --pairs 8 --chunk 8 $g --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0 --steps 10 --warmup 3 > $O/bench_p8$g.json 2> $O/bench_p8$g.err; echo "bench8 $g rc=$?"; tail -2 $O/bench_p8$g.err
python - <<PY
import json
d=json.load(open("$O/bench_p8$g.json"))
print("pairs8 $g value %.1f e2e %.1f ms/step %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
PY
done
