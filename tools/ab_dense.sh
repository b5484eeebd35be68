#!/bin/bash
# A/B build variants of the dense integrate kernel on the GPU box (nvcc is in the image).
# usage: tools/ab_dense.sh "<flags variant 1>" "<flags variant 2>" ...
for v in "$@"; do
  TSDF_NVCC_EXTRA="$v" python -m union_thesis_slam_b200.build --force 2>&1 | grep -A3 "k_integrate_denseILb1" | grep -E "registers|spill" | tr '\n' ' '
  echo; echo "== $v"
  python tools/quick_dense.py 0.01 4 3 2>&1 | grep "^frame" | tail -1
  python tools/quick_dense.py 0.01 4 40 2>&1 | grep "^frame" | tail -1
done
