#!/bin/bash
# Bench sweeps: engine chunk size and CUDA graphs (one JSON line each, trimmed).
set -u
O=gpurun_out/${1:-s10}
mkdir -p $O
for c in 11 16 22 32; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --latency-frames 0 --chunk $c > $O/bench_chunk$c.json 2> $O/bench_chunk$c.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_chunk$c.json")); print("chunk $c eager : value %.1f e2e %.1f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e: print("chunk $c failed", e)
PY
done
for c in 11 22; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --latency-frames 0 --chunk $c --graphs > $O/bench_chunk${c}_graphs.json 2> $O/bench_chunk${c}_graphs.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_chunk${c}_graphs.json")); print("chunk $c graphs: value %.1f e2e %.1f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e: print("chunk $c graphs failed", e)
PY
done
tail -3 $O/*.err
