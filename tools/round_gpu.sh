#!/bin/bash
# One-GPU round-end job (run under gpurun): GPU tests, smoke, bench (both arms), ncu launch list of the
# bench command, one `--set full` capture of the dense integrate kernel.  Outputs land in gpurun_out/.
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r1c}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f frac %.3f serial %.3f e2e %.4g hash %.4g hash_frac %.3f cpu %.3g" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["serialised"]["frac"], d["e2e"]["value"], d["hash"].get("value",0), d["hash"].get("roofline",{}).get("frac",0), d.get("cpu_baseline",{}).get("value",0)))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 24 --warmup 3 --no-cpu > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate_dense -s 80 -c 2 -f -o $O/${TAG}_dense_full \
    python tools/quick_dense.py 0.01 16 3 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/quick_classes.py 8 1 2>&1 | grep world > $O/${TAG}_classes.log; cat $O/${TAG}_classes.log
ls -la $O | tail -n 12
