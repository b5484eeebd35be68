#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2k}
( timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_hash_gpu.py -m gpu -x -q ) > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -n 4 $O/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f frac %.3f serial %.3f e2e %.4g (raw %.4g, blocking %.4g)" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["serialised"]["frac"], d["e2e"]["value"], d["e2e"]["raw_u16"]["value"], d["e2e"]["blocking_read"]["value"]))
h=d.get("hash",{}); print("hash", h.get("value"), h.get("roofline",{}).get("frac"), "e2e", h.get("e2e",{}).get("value"), h.get("error"))
PY
