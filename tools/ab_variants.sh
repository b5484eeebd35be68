#!/bin/bash
# Build compile-time variants of the library HERE (CPU container), then measure them in one gpurun call:
#   bash tools/ab_variants.sh build        # writes union_thesis_slam_b200/variants/lib_<name>.so
#   gpurun -- 'bash tools/ab_variants.sh run'
set -u
V=union_thesis_slam_b200/variants
declare -A FLAGS=( [base]="" [evictlast]="-DPIXEL_EVICT_LAST=1" [unroll]="-DQUAD_UNROLL=1" [l2pf]="-DINTEGRATE_L2_PREFETCH=1" )
if [ "${1:-}" = "build" ]; then
  mkdir -p $V
  for n in "${!FLAGS[@]}"; do
    TSDF_LIB_OUT=$V/lib_$n.so TSDF_NVCC_EXTRA="${FLAGS[$n]}" python -m union_thesis_slam_b200.build --force 2>&1 | grep -E "k_integrate_denseILb1|registers|error" | grep -A1 "k_integrate_denseILb1" | tail -n 1 | sed "s/^/$n: /"
  done
else
  for n in base evictlast unroll l2pf base; do
    echo "== $n"; TSDF_B200_LIB=$PWD/$V/lib_$n.so python tools/quick_classes.py 1 2>&1 | grep "classes 2" | tail -n 1
  done
fi
