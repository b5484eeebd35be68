#!/bin/bash
# A/B of compile-time variants (lib_var_*.so beside the product library): event timing + ncu instruction counts
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2i}
L=$PWD/union_thesis_slam_b200
for lib in $L/libtsdf_b200.so $L/lib_var_*.so; do
  echo "=== $(basename $lib)"
  TSDF_B200_LIB=$lib timeout 300 python tools/quick_batch.py 1 2>/dev/null | head -2
  TSDF_B200_LIB=$lib timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_integrate_dense -s 80 -c 1 python tools/quick_dense.py 0.01 16 3 2>&1 | grep -E "inst_executed|duration|issue_active"
done > $O/${TAG}_variants.log 2>&1
cat $O/${TAG}_variants.log
