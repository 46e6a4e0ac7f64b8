#!/bin/bash
# debug build (see tools/r2_session_u.sh): GRU candidate convolutions without the fp32 store of the state (RPE_CONV_DEBUG=8)
set -u
O=gpurun_out/r2_u2
mkdir -p $O
cp robust-pose-estimator_b200/librpe_b200.so /tmp/librpe_prod.so
cp robust-pose-estimator_b200/build/librpe_b200_dbg.so robust-pose-estimator_b200/librpe_b200.so
for d in 0 8 0 8; do
  echo "== RPE_CONV_DEBUG=$d" >> $O/conv_probe_state_store.txt
  RPE_CONV_DEBUG=$d timeout 300 python tools/conv_probe.py --n=64 --only=q1 --only=q2 >> $O/conv_probe_state_store.txt 2>&1
done
cp /tmp/librpe_prod.so robust-pose-estimator_b200/librpe_b200.so
cat $O/conv_probe_state_store.txt
