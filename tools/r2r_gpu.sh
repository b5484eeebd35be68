#!/bin/bash
set -u
L=$PWD/union_thesis_slam_b200
for lib in $L/libtsdf_b200.so $L/lib_var_*.so $L/libtsdf_b200.so; do
  echo "=== $(basename $lib)"
  TSDF_B200_LIB=$lib timeout 300 python tools/quick_batch.py 1 2>/dev/null | head -2
  TSDF_B200_LIB=$lib timeout 300 python tools/quick_batch.py 1 8 3 2>/dev/null | head -1
  TSDF_B200_LIB=$lib timeout 300 python tools/quick_batch.py 1 8 0 2>/dev/null | head -1
done
