#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2b}
( timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_dense_gpu.py tests/test_long_history_gpu.py -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 $O/${TAG}_pytest.log
( TSDF_DENSE_V1=1 timeout 300 python tools/quick_batch.py 1; for b in 1 2 4; do timeout 300 python tools/quick_batch.py $b; done; for b in 1 4; do timeout 300 python tools/quick_batch.py $b 8 0; done ) 2>/dev/null > $O/${TAG}_ab.log; cat $O/${TAG}_ab.log
