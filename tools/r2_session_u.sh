#!/bin/bash
# debug-switch probe of the convolution kernel (library built with -DRPE_CONV_DEBUG_BUILD, swapped in on the GPU box only).
# Build it first, in the container:
#   cd robust-pose-estimator_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
#        -DRPE_CONV_DEBUG_BUILD -c csrc/conv.cu -o /tmp/conv_dbg.o && \
#   nvcc -shared -o build/librpe_b200_dbg.so /tmp/conv_dbg.o $(ls build/*.o | grep -v build/conv.o) -gencode arch=compute_100a,code=sm_100a
set -u
O=gpurun_out/r2_u
mkdir -p $O
cp robust-pose-estimator_b200/librpe_b200.so /tmp/librpe_prod.so
cp robust-pose-estimator_b200/build/librpe_b200_dbg.so robust-pose-estimator_b200/librpe_b200.so
for d in 0 1 2 4; do
  echo "== RPE_CONV_DEBUG=$d (1 no loads, 2 no MMAs, 4 no epilogue memory traffic)" >> $O/conv_probe_debug.txt
  RPE_CONV_DEBUG=$d timeout 300 python tools/conv_probe.py --n=32 >> $O/conv_probe_debug.txt 2>&1
done
cp /tmp/librpe_prod.so robust-pose-estimator_b200/librpe_b200.so
cat $O/conv_probe_debug.txt
