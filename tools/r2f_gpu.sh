#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2f}
L=$PWD/union_thesis_slam_b200
( timeout 900 python -m pytest tests/test_batch_gpu.py -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest.log
( echo "== v1 sunk loads 64 regs"; TSDF_DENSE_V1=1 timeout 300 python tools/quick_batch.py 1;
  for v in early64 early72 early80; do echo "== v1 $v"; TSDF_B200_LIB=$L/lib_$v.so TSDF_DENSE_V1=1 timeout 300 python tools/quick_batch.py 1; done
  echo "== v2c"; for b in 1 2 4; do timeout 300 python tools/quick_batch.py $b; done;
  echo "== world 8"; TSDF_DENSE_V1=1 timeout 300 python tools/quick_batch.py 1 8 0; TSDF_B200_LIB=$L/lib_early80.so TSDF_DENSE_V1=1 timeout 300 python tools/quick_batch.py 1 8 0; for b in 1 4; do timeout 300 python tools/quick_batch.py $b 8 0; done ) 2>/dev/null > $O/${TAG}_ab.log; cat $O/${TAG}_ab.log
