#!/bin/bash
# 8-GPU box: world-8 parity check, then the dense bench at N = 1, 2, 4, 8 (add "hash" as $1 for the hash map at 1 and 8).
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tests/multi_gpu_check.py > $O/s8_check.log 2>&1; echo "check rc=$?"; tail -n 1 $O/s8_check.log
timeout 300 python bench.py --no-hash --no-cpu > $O/s8_dense_n1.json 2> $O/s8_err.log; echo "n1 rc=$?"
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29520 + n)) bench.py --gpus $n > $O/s8_dense_n$n.json 2>> $O/s8_err.log; echo "n$n rc=$?"
done
if [ "${1:-}" = "hash" ]; then
  timeout 300 python bench.py --workload hash --steps 128 --warmup 32 > $O/s8_hash_n1.json 2>> $O/s8_err.log; echo "h1 rc=$?"
  timeout 300 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --workload hash --steps 128 --warmup 32 > $O/s8_hash_n8.json 2>> $O/s8_err.log; echo "h8 rc=$?"
fi
python - <<PY
import json, os
def rd(f):
    try: return json.loads([l for l in open("$O/"+f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e: return None
base = rd("s8_dense_n1.json")
for n in (1, 2, 4, 8):
    x = rd("s8_dense_n%d.json" % n)
    if x: print("dense n%d value %.4g ms/step %.4f frac %.3f e2e %.4g  x%.2f" % (n, x["value"], x["ms_per_step"], x["roofline"]["frac"], x["e2e"]["value"], x["value"] / base["value"]))
PY
