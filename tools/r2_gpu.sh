#!/bin/bash
# Round-2 one-GPU job (run under gpurun): GPU tests, smoke, bench (both arms).  usage: tools/r2_gpu.sh TAG [skip-tests]
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2a}
ls oracle/_ref > $O/${TAG}_ref_ls.log 2>&1
if [ "${2:-}" != "skip-tests" ]; then
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $O/${TAG}_smoke.log
fi
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -n 4 $O/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f frac %.3f serial %.3f e2e %.4g (raw %.4g, blocking %.4g)" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["serialised"]["frac"], d["e2e"]["value"], d["e2e"]["raw_u16"]["value"], d["e2e"]["blocking_read"]["value"]))
h=d.get("hash",{}); print("hash", h.get("value"), h.get("roofline",{}).get("frac"), "e2e", h.get("e2e",{}).get("value"), h.get("error"))
for k,v in d.get("configs",{}).items():
    print(k, v if isinstance(v,str) else "%.4g upd/s %.4f ms/frame frac %.3f" % (v["value"], v["ms_per_step"], v["roofline"]["frac"]))
print("clocks", d.get("clocks")); print("cpu", json.dumps(d.get("cpu_baseline"))[:600]); print("hash cpu", json.dumps(h.get("cpu_baseline"))[:500])
PY
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; echo "ref rc=$?"; cut -c1-300 $O/${TAG}_bench_reference.json; tail -n 3 $O/${TAG}_bench.err
