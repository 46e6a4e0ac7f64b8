#!/bin/bash
# Round 2, GPU session P (8 GPUs): the bench at N = 8 and N = 4 exactly as the driver launches it
set -u
O=gpurun_out/r2_p
mkdir -p $O
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench n$N rc=$?"; grep -v "^$\|\*\*\*\|OMP_NUM" $O/bench_n$N.err | tail -4
python - <<PY
import json
try:
    d=json.load(open("$O/bench_n$N.json"))
    print("N=%d value %.1f e2e %.1f ms/step %.1f scaling %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"]))
    print("strong", d.get("strong")); c=d["config5"]; print("config5 pairs/s %.1f seconds %.3f failed %d ate %s" % (c["pairs_per_s"], c["seconds"], c["failed_pairs"], c["accuracy"]["ate_rmse_mm"])); print("parity", d["pose_parity"]["max_rel_translation"], "clocks", d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
done
