#!/bin/bash
# Round 2, GPU session B: fp16 split planes + batch-invariant pose solver: whole GPU suite + bench
set -u
O=gpurun_out/r2_b
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 600 -x --deselect tests/test_gpu_configs.py::test_sharded_halo_run_is_bit_equal_to_single_process > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "config1|bench64|sharded|only3d|passed|failed|FAILED|corr fp16x3" $O/pytest_gpu.log | head -60
timeout 900 python bench.py --no-cpu-baseline --config5-frames 0 --latency-pairs 40 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench.json"))
    print("value %.1f e2e %.1f ms/step %.1f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
    print("parity", d["pose_parity"]); print("gpu_reference", d.get("gpu_reference",{}).get("value"))
    for k,v in d["stages"].items(): print("%-18s launches %5d total_ms %9.2f avg_us %9.1f" % (k, v["launches"], v["total_ms"], v["avg_us"]))
    for k,v in d["kernels"].items(): print("%-26s %s achieved %8.1f frac %.3f" % (k, v["unit"], v["achieved"], v["frac"]))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -s --timeout 600 -k sharded 2>&1 | grep -E "sharded|passed|failed|Error" | head
