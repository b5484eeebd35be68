#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2g}
( timeout 1200 python -m pytest tests/test_batch_gpu.py tests/test_dense_gpu.py tests/test_hash_gpu.py tests/test_long_history_gpu.py tests/test_hd_stream_gpu.py tests/test_fuzz_gpu.py -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/${TAG}_pytest.log
( timeout 300 python tools/quick_batch.py 1; timeout 300 python tools/quick_batch.py 4; timeout 300 python tools/quick_batch.py 1 8 0 ) 2>/dev/null > $O/${TAG}_ab.log; cat $O/${TAG}_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 80 -c 2 -f -o $O/${TAG}_v1 python tools/quick_dense.py 0.01 16 3 > $O/${TAG}_ncu_v1.log 2>&1; echo "ncu v1 rc=$?"
