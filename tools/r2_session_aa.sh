#!/bin/bash
set -u
O=gpurun_out/r2_aa
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0"
timeout 600 ncu --set full --clock-control none -k regex:conv_f16x3_pair -s 97 -c 1 -o $O/fh1_full $B --steps 1 --warmup 0 > /dev/null 2> $O/fh1_full.err; echo "fh1 rc=$?"
ncu -i $O/fh1_full.ncu-rep --page raw --csv > $O/fh1_full_raw.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("$O/fh1_full_raw.csv")))
h=rows[0]; r=rows[2]
for k in ("Kernel Name","gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","launch__registers_per_thread"):
    print(k, r[h.index(k)][:80])
PY
rm -f $O/fh1_full.ncu-rep
timeout 900 $B > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print("value %.1f e2e %.1f ms/step %.1f clocks %s parity %.3e" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["pose_parity"]["max_rel_translation"]))
for k in ("conv_tc","conv_tc_enc","corr_build","pose_solve"): v=d["stages"][k]; print("%-18s total_ms %9.2f avg_us %9.1f" % (k, v["total_ms"], v["avg_us"]))
PY
