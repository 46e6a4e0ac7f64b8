#!/bin/bash
# One GPU call: parity tests, bench lines (both arms), conv probe with the debug switches, ncu --set full of the hand-written kernels.
# Only text / csv summaries are kept (gpurun_out/ is capped at 64 MiB): .ncu-rep files are converted and deleted on the box.
set -u
O=gpurun_out/${1:-s5}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rs -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for d in 0 1 2 4; do
  RPE_CONV_DEBUG=$d timeout 300 python tools/conv_probe.py --n=22 > $O/probe_dbg$d.txt 2>&1
done
RPE_CONV_PAIR=0 timeout 300 python tools/conv_probe.py --n=22 > $O/probe_nopair.txt 2>&1
T=/tmp/ncu_tmp; mkdir -p $T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_bf16 -o $T/conv -f \
  python tools/conv_probe.py --n=22 --once --only=convc1 --only=convc2 --only=conv --only=zr1 --only=q1 --only=fh1 --only="enc l1" --only="enc l2a" > $O/ncu_conv.log 2>&1
ncu -i $T/conv.ncu-rep --page raw --csv > $O/ncu_conv_probe_n22_raw.csv 2>/dev/null
ncu -i $T/conv.ncu-rep --page details --csv > $O/ncu_conv_probe_n22_details.csv 2>/dev/null
ncu -i $T/conv.ncu-rep --page source --csv --kernel-name regex:conv_bf16_pair --launch-skip 3 --launch-count 1 > $O/ncu_conv_zr1_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none \
  -k regex:'corr_lookup_nhwc|pose_solve|corr_gemm|corr_pool|norm_act|instnorm_partial|im2col7s2|convex_upsample8|warp8_mask|depth_proj' \
  -s 30 -c 24 -o $T/own -f \
  python bench.py --steps 1 --warmup 1 --pairs 11 --no-cpu-baseline --latency-frames 0 > $O/ncu_own.log 2>&1
ncu -i $T/own.ncu-rep --page raw --csv > $O/ncu_own_bench_pairs11_raw.csv 2>/dev/null
ncu -i $T/own.ncu-rep --page details --csv > $O/ncu_own_bench_pairs11_details.csv 2>/dev/null
gzip -f $O/*_details.csv $O/ncu_conv_zr1_source.csv
du -sh $O; ls -la $O
