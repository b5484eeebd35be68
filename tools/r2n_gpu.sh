#!/bin/bash
set -u
L=$PWD/union_thesis_slam_b200
python -m pytest tests/test_bucket_gpu.py tests/test_hash_gpu.py -m gpu -x -q 2>&1 | tail -1
for lib in $L/libtsdf_b200.so $L/lib_var_*.so; do
  echo "=== $(basename $lib)"
  TSDF_B200_LIB=$lib timeout 300 python bench.py --workload hash --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hash value %.4g ms/step %.4f frac %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))"
done
