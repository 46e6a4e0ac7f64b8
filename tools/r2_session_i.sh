#!/bin/bash
# Round 2, GPU session I (2 GPUs): the bench at N = 2 exactly as the driver launches it (weak primary + strong leg + config 5 sharded)
set -u
O=gpurun_out/r2_i
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -5 $O/bench_n2.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_n2.json"))
    print("N=%d value %.1f e2e %.1f ms/step %.1f scaling %s launches %d" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"], d["gpu_launches"]))
    print("strong", d.get("strong")); print("config5", d["config5"]); print("parity", d["pose_parity"]); print("clocks", d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/ref_n2.json 2> $O/ref_n2.err; echo "ref n2 rc=$?"; cut -c1-300 $O/ref_n2.json
