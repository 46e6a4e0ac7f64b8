#!/bin/bash
# pose solver built with __launch_bounds__(256, 4) and the 2-CTAs/SM cap lifted (csrc/pose.cu edited by sed into /tmp, linked into
# build/librpe_b200_pose4.so, swapped in on the GPU box only): 64 registers, spills in the pixel loop -> slower (profiles/r2_z_*).
set -u
O=gpurun_out/r2_z2
mkdir -p $O
cp robust-pose-estimator_b200/librpe_b200.so /tmp/librpe_prod.so
cp robust-pose-estimator_b200/build/librpe_b200_pose4.so robust-pose-estimator_b200/librpe_b200.so
timeout 600 python tools/pose_probe.py 32 > $O/pose_probe_4ctas.txt 2>&1; cat $O/pose_probe_4ctas.txt
cp /tmp/librpe_prod.so robust-pose-estimator_b200/librpe_b200.so
