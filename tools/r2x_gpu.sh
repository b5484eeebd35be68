#!/bin/bash
set -u
for m in f64 u16; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/quick_fan2.py $m 2>&1 | grep -v -i "warn\|Voxel\|^$\|\*\*\*"
done
