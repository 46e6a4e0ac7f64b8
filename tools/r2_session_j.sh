#!/bin/bash
set -u
O=gpurun_out/r2_j
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_corr.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 python tools/corr_probe.py 64 2>&1 | grep -v Warn | tee $O/corr_probe_residentA.txt
RPE_CORR_STREAM_A=1 timeout 300 python tools/corr_probe.py 64 2>&1 | grep -v Warn | grep planes | tee $O/corr_probe_streamA.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --kernel-name-base demangled -k regex:conv_f16x3_pair_kernel -c 6 --csv --log-file $O/corr_ncu_metrics.csv python tools/corr_probe.py 64 > /dev/null 2>&1
grep -v "^==" $O/corr_ncu_metrics.csv | tail -40 | cut -d, -f5,13-15 | tail -36
