#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2d}
TSDF_DENSE_V1=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 80 -c 2 -f -o $O/${TAG}_v1 python tools/quick_dense.py 0.01 16 3 > $O/${TAG}_ncu_v1.log 2>&1; echo "ncu v1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 80 -c 2 -f -o $O/${TAG}_v2 python tools/quick_dense.py 0.01 16 3 > $O/${TAG}_ncu_v2.log 2>&1; echo "ncu v2 rc=$?"
TSDF_DENSE_V1=1 timeout 300 python tools/quick_dense.py 0.01 16 3 2>&1 | grep rep
timeout 300 python tools/quick_dense.py 0.01 16 3 2>&1 | grep rep
