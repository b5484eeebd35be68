#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2c}
( timeout 900 python -m pytest tests/test_batch_gpu.py -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/${TAG}_pytest.log
for b in 1 4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 60 -c 2 -f -o $O/${TAG}_v2_b$b python tools/quick_batch.py $b > $O/${TAG}_ncu_b$b.log 2>&1; echo "ncu b$b rc=$?"
done
