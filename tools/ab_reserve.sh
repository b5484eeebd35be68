#!/bin/bash
for v in "$@"; do
  TSDF_NVCC_EXTRA="$v" python -m union_thesis_slam_b200.build --force >/dev/null 2>&1
  echo "== $v"
  python bench.py --no-hash --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['serialised']['frac'])"
  python tools/quick_shard.py 8 2>&1 | tail -1
done
