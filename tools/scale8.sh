#!/bin/bash
# 8-GPU box: world-8 parity check, then the bench at N=1 and N=8 (dense), N=1 and N=8 (hash).
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tests/multi_gpu_check.py > $O/s8_check.log 2>&1; echo "check rc=$?"; tail -n 3 $O/s8_check.log
timeout 300 python bench.py --steps 512 --warmup 32 --no-hash --no-cpu > $O/s8_dense_n1.json 2> $O/s8_err.log; echo "n1 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 512 --warmup 32 > $O/s8_dense_n8.json 2>> $O/s8_err.log; echo "n8 rc=$?"
timeout 300 python bench.py --workload hash --steps 128 --warmup 32 > $O/s8_hash_n1.json 2>> $O/s8_err.log; echo "h1 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --workload hash --steps 128 --warmup 32 > $O/s8_hash_n8.json 2>> $O/s8_err.log; echo "h8 rc=$?"
python - <<PY
import json
def rd(f):
    try: return json.loads([l for l in open("$O/"+f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e: return {"error": repr(e)}
a,b,c,d=rd("s8_dense_n1.json"),rd("s8_dense_n8.json"),rd("s8_hash_n1.json"),rd("s8_hash_n8.json")
for n,x in (("dense n1",a),("dense n8",b),("hash n1",c),("hash n8",d)):
    print(n, {k:x.get(k) for k in ("value","ms_per_step","error")}, (x.get("roofline") or {}).get("frac"), (x.get("e2e") or {}).get("value"))
try: print("dense x%.2f hash x%.2f" % (b["value"]/a["value"], d["value"]/c["value"]))
except Exception as e: print(e)
PY
