"""Multi-GPU parity as a collected test: tests/multi_gpu_check.py under torchrun, one rank per GPU (keyframe-sharded pass +
ncclAllReduce of the packed blocks against the oracle, the compact shared-landmark exchange over three consecutive rounds,
disjoint shards, and the distributed solve against the single-GPU solve with x identical on every rank).
Skipped when the box has fewer GPUs than ranks; the three-rank case is there because a shared landmark that a rank never
observes only exists from three ranks on (the round-1 advisor finding)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_multi_gpu_check_under_torchrun(world):
    n = _gpu_count()
    if n < world:
        pytest.skip("needs %d GPUs on the box, found %d" % (world, n))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    assert out.stdout.count("multi_gpu_check OK") >= 4, out.stdout[-1500:]
