#!/bin/bash
# Round-2 one-GPU job (run under gpurun): GPU tests, bench (both arms), ncu instruction counts and one
# `--set full` capture of the dense integrate kernel.  Outputs land in gpurun_out/.  usage: tools/r2_gpu.sh TAG [skip-tests]
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2a}
if [ "${2:-}" != "skip-tests" ]; then
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/${TAG}_pytest_gpu.log
fi
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f frac %.3f serial %.3f e2e %.4g hash %.4g hash_frac %.3f cpu %.3g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["serialised"]["frac"], d["e2e"]["value"], d.get("hash",{}).get("value",0), d.get("hash",{}).get("roofline",{}).get("frac",0), d.get("cpu_baseline",{}).get("value",0)))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"; tail -c 600 $O/${TAG}_bench_reference.json
timeout 600 python tools/quick_dense.py 0.01 16 3 2>&1 | grep rep > $O/${TAG}_quick_dense.log; cat $O/${TAG}_quick_dense.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 80 -c 2 -f -o $O/${TAG}_dense_full \
    python tools/quick_dense.py 0.01 16 3 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -n 8
