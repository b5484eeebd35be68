#!/bin/bash
# One-GPU verification after the fan-out work: GPU tests, smoke, both bench arms (the driver's command).
set -u
O=gpurun_out; TAG=r2y
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/${TAG}_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f frac %.3f e2e %.4g u16 %.4g blocking %.4g hash %.4g hash_frac %.3f hash_e2e %.4g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["raw_u16"]["value"], d["e2e"]["blocking_read"]["value"], d["hash"].get("value",0), d["hash"].get("roofline",{}).get("frac",0), d["hash"]["e2e"]["value"]))
r=json.loads(open("$O/${TAG}_bench_reference.json").read().strip().splitlines()[-1])
print("reference %.4g %s" % (r["value"], r["cpu_baseline"]["kind"]))
PY
