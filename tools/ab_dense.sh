#!/bin/bash
# A/B build variants of the dense integrate kernel on the GPU box (nvcc is in the image).
for v in "-DINTEGRATE_STAGED=1 -DINTEGRATE_MIN_CTAS=12" "-DINTEGRATE_STAGED=0 -DINTEGRATE_MIN_CTAS=12" "-DINTEGRATE_STAGED=0 -DINTEGRATE_MIN_CTAS=10" "-DINTEGRATE_STAGED=1 -DINTEGRATE_MIN_CTAS=10" "-DINTEGRATE_STAGED=0 -DINTEGRATE_MIN_CTAS=8"; do
  TSDF_NVCC_EXTRA="$v" python -m union_thesis_slam_b200.build --force 2>&1 | grep -E "k_integrate_denseILb1|error" -A2 | grep -E "registers|error" | head -1
  echo "== $v"; python tools/quick_dense.py 0.01 16 2>&1 | grep "^rep 2"
done
