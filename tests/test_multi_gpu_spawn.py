"""Multi-process NCCL parity, collected by pytest (-m gpu): spawns tests/multi_gpu_check.py under
torch.distributed.run on every GPU of the box (dense x-sheets and owner-hashed blocks gathered over
NCCL must equal the CPU oracle bit for bit).  On a one-GPU box the same check runs with world 1,
so the gather code path is exercised either way."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_torchrun_dense_and_hash_gather_equals_oracle():
    import torch
    world = min(torch.cuda.device_count(), 8)
    assert world >= 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    res = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-4000:]
    assert "multi_gpu_check world=%d" % world in res.stdout and "-> OK" in res.stdout, res.stdout[-4000:]
