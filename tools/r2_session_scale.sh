#!/bin/bash
# The driver's scaling run on ONE 8-GPU box: N = 1, 2, 4, 8 back to back, launched the way the driver launches them.
set -u
O=gpurun_out/r2_scale
mkdir -p $O
LIGHT="--no-cpu-baseline --no-gpu-reference --latency-pairs 0"
for N in 1 2 4 8; do
if [ $N = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 $LIGHT > $O/bench_n$N.json 2> $O/bench_n$N.err
else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 5 --warmup 3 $LIGHT > $O/bench_n$N.json 2> $O/bench_n$N.err; fi
echo "bench n$N rc=$?"; grep -v "^$\|\*\*\*\|OMP_NUM" $O/bench_n$N.err | tail -3
python - <<PY
import json
try:
    d=json.load(open("$O/bench_n$N.json"))
    print("N=%d value %.1f e2e %.1f ms/step %.1f scaling %s clocks %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"], d["clocks"]["sm_mhz"]))
    if d.get("strong"): print("  strong %.1f pairs/s (%.2f ms)" % (d["strong"]["value"], d["strong"]["ms_per_step"]))
    c=d.get("config5")
    if c: print("  config5 pairs/s %.1f seconds %.3f failed %d ate %.3f" % (c["pairs_per_s"], c["seconds"], c["failed_pairs"], c["accuracy"]["ate_rmse_mm"]))
    print("  parity", d["pose_parity"]["max_rel_translation"])
except Exception as e:
    print("parse failed", e)
PY
done
