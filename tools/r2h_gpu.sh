#!/bin/bash
# quick kernel iteration: parity subset, event timing, ncu instruction counts.  usage: tools/r2h_gpu.sh TAG
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2h}
( timeout 1200 python -m pytest tests/test_dense_gpu.py tests/test_hash_gpu.py tests/test_long_history_gpu.py -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest.log
( timeout 300 python tools/quick_batch.py 1; timeout 300 python tools/quick_batch.py 1 8 0 ) 2>/dev/null > $O/${TAG}_ab.log; cat $O/${TAG}_ab.log
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_integrate_dense -s 80 -c 2 python tools/quick_dense.py 0.01 16 3 2>&1 | grep -E "k_integrate|inst_executed|duration|issue_active" > $O/${TAG}_ncu_counts.log; cat $O/${TAG}_ncu_counts.log
