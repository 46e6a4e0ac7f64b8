#!/bin/bash
# Round 2, GPU session H: ncu evidence.  (1) launch list + DRAM bytes of every kernel of the default bench configuration,
# (2) --set full captures of the kernels the judge named: fused correlation volume (tensor pipe), GRU conv, lookup, pose solver, norm_act.
set -u
O=gpurun_out/r2_h
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 9000 --csv \
    --log-file $O/launches_traffic.csv $B --steps 2 --warmup 1 > $O/bench_under_ncu.json 2> $O/launches.err; echo "launch list rc=$?"
wc -l $O/launches_traffic.csv
S="$B --steps 1 --warmup 0"
cap() { # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 -o $O/$1 $S > /dev/null 2> $O/$1.err; echo "$1 rc=$?"
}
cap corr_volume 'conv_f16x3_pair_kernel<7>' 0 1
cap conv_gates 'conv_f16x3_pair_kernel<1>' 4 1
cap conv_state 'conv_f16x3_pair_kernel<2>' 4 1
cap conv_planes 'conv_f16x3_pair_kernel<4>' 60 2
cap conv_f32stats 'conv_f16x3_pair_kernel<6>' 3 2
cap lookup 'corr_lookup_nhwc_r4' 3 1
cap pose_solve 'pose_solve_kernel' 0 1
cap norm_act 'norm_act_kernel' 2 1
ls -la $O
