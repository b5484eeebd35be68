#!/bin/bash
# multi-GPU job (gpurun --gpus N): torchrun parity check + bench at N.  usage: tools/r2_multi.sh TAG N
set -u
O=gpurun_out; mkdir -p $O
TAG=${1:-r2m}; N=${2:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $O/${TAG}_check_n$N.log 2>&1; echo "check rc=$?"; grep multi_gpu_check $O/${TAG}_check_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 ) > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "bench rc=$?"; tail -n 5 $O/${TAG}_bench_n$N.err
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.4g ms/step %.4f frac %.3f e2e %.4g (raw %.4g, blocking %.4g)" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["raw_u16"]["value"], d["e2e"]["blocking_read"]["value"]))
h=d.get("hash",{}); print("hash", h.get("value"), h.get("roofline",{}).get("frac"), "e2e", h.get("e2e",{}).get("value"), h.get("error"))
for k,v in d.get("configs",{}).items():
    print(k, v if isinstance(v,str) else "%.4g upd/s %.4f ms/frame frac %.3f" % (v["value"], v["ms_per_step"], v["roofline"]["frac"]))
PY
[ "${3:-}" = "noref" ] || ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) 2>&1 | grep -E "impl|real" | cut -c1-400
