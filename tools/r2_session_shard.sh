#!/bin/bash
set -u
O=gpurun_out/r2_shard
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/sharded_edge_probe.py > $O/sharded_edge.txt 2>&1; echo "rc=$?"
grep -v "^$\|\*\*\*\|OMP_NUM" $O/sharded_edge.txt | tail -14
