#!/bin/bash
# 8 GPUs: the torchrun parity check (incl. fan-out) and a short N=8 bench (no configs / cpu).
set -u
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --steps 20 --warmup 5 --no-configs > gpurun_out/r2u_bench_n8.json 2> gpurun_out/r2u_bench_n8.err
tail -3 gpurun_out/r2u_bench_n8.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2u_bench_n8.json").read().strip().splitlines()[-1])
    e = j["e2e"]
    print("value %.4g  e2e %.4g  u16 %.4g  blocking %.4g  per_rank %.4g" % (j["value"], e["value"], e["raw_u16"]["value"], e["blocking_read"]["value"], e.get("per_rank_upload", {}).get("value", 0)))
    h = j["hash"]
    print("hash value %.4g e2e %.4g" % (h["value"], h["e2e"]["value"]))
except Exception as exc:
    print("parse failed", exc)
PY
