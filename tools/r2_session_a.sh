#!/bin/bash
# Round 2, GPU session A: new correlation kernel + new config tests first, then the whole GPU suite, smoke and the default bench line.
set -u
O=gpurun_out/r2_a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_corr.py -m gpu -q -x -s --timeout 300 > $O/pytest_corr.log 2>&1; echo "corr rc=$?"
tail -15 $O/pytest_corr.log
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -s --timeout 600 > $O/pytest_configs.log 2>&1; echo "configs rc=$?"
tail -25 $O/pytest_configs.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_corr.py --deselect tests/test_gpu_configs.py > $O/pytest_rest.log 2>&1; echo "rest rc=$?"
tail -8 $O/pytest_rest.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -5 $O/bench_default.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_default.json"))
    print("value %.1f e2e %.1f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
    print("latency", d["latency"])
    print("cpu_baseline", d.get("cpu_baseline")); print("gpu_reference", d.get("gpu_reference"))
    print("roofline", d["roofline"]["frac"], d["roofline"]["executed_frac"], "clocks", d["clocks"])
    print("parity", d["pose_parity"]); print("config5", d["config5"])
    for k,v in d["stages"].items(): print("%-18s launches %5d total_ms %9.2f avg_us %9.1f" % (k, v["launches"], v["total_ms"], v["avg_us"]))
    for k,v in d["kernels"].items(): print("%-26s %s achieved %8.1f frac %.3f" % (k, v["unit"], v["achieved"], v["frac"]))
except Exception as e:
    print("bench parse failed", e)
PY
