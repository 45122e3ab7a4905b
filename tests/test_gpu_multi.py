"""Sharded (multi-GPU) parity: needs >= 2 CUDA devices, otherwise skipped (CPU coverage: tests/test_sharding.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["nccl", "peer_memory"])
def test_two_rank_solve_matches_oracle(exchange):
    """exchange = peer_memory: the compact exchange of the owner-computes layout through CUDA-IPC-mapped peer buffers (GLBA_P2P=1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 1000 + (1 if exchange == "peer_memory" else 0)
    env = dict(os.environ, GLBA_P2P="1" if exchange == "peer_memory" else "0")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["ok"] and res["world"] == 2
