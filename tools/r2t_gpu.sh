#!/bin/bash
# 2 GPUs: new fan-out tests, the torchrun parity check, and a short N=2 bench (no configs / cpu).
set -u
timeout 900 python -m pytest tests/test_multi_gpu_spawn.py -x -q -m gpu > gpurun_out/r2t_pytest.log 2>&1
tail -5 gpurun_out/r2t_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 20 --warmup 5 --no-configs > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
tail -3 gpurun_out/r2t_bench_n2.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2t_bench_n2.json").read().strip().splitlines()[-1])
    e = j["e2e"]
    print("value %.4g  e2e %.4g  u16 %.4g  blocking %.4g  per_rank %.4g" % (j["value"], e["value"], e["raw_u16"]["value"], e["blocking_read"]["value"], e.get("per_rank_upload", {}).get("value", 0)))
    h = j["hash"]
    print("hash value %.4g e2e %.4g" % (h["value"], h["e2e"]["value"]))
except Exception as exc:
    print("parse failed", exc)
PY
