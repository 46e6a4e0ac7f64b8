#!/bin/bash
# Round 2, GPU session X: ncu evidence on the final tree.  (1) launch list + DRAM bytes of every kernel of the default bench
# configuration, (2) --set full capture of 14 consecutive convolution launches inside the update operator (all epilogue kinds).
set -u
O=gpurun_out/r2_x
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-gpu-reference --config5-frames 0 --latency-pairs 0"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 9000 --csv \
    --log-file $O/launches_traffic.csv $B --steps 2 --warmup 1 > $O/bench_under_ncu.json 2> $O/launches.err; echo "launch list rc=$?"
wc -l $O/launches_traffic.csv
gzip -f $O/launches_traffic.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_f16x3_pair -s 90 -c 14 -o $O/conv_full $B --steps 1 --warmup 0 > /dev/null 2> $O/conv_full.err; echo "conv_full rc=$?"
tail -3 $O/conv_full.err
ls -la $O
