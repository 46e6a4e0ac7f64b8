#!/bin/bash
# Profiling call: torch-profiler step breakdown + ncu --set full (with SASS-level source pages) of the conv probe.
set -u
O=gpurun_out/${1:-s8}
mkdir -p $O
timeout 300 python tools/step_profile.py --pairs 22 --chunk 11 > $O/step_profile_pairs22.txt 2>&1
T=/tmp/ncu_tmp; mkdir -p $T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_bf16 -o $T/conv -f \
  python tools/conv_probe.py --n=22 --once --only=convc1 --only=zr1 --only=q1 --only="enc stem" --only="enc l2 " --only=convf2 > $O/ncu_conv.log 2>&1
ncu -i $T/conv.ncu-rep --page raw --csv > $O/ncu_conv_probe_n22_raw.csv 2>/dev/null
i=0
for name in convc1 convf2 zr1 q1 enc_stem enc_l2; do
  ncu -i $T/conv.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > $O/ncu_src_$name.csv 2>/dev/null
  i=$((i+1))
done
gzip -f $O/ncu_src_*.csv $O/ncu_conv_probe_n22_raw.csv
head -30 $O/step_profile_pairs22.txt; ls -la $O
