"""Multi-GPU exchange tests (need >= 2 B200s; skipped on a single-GPU box).  Two processes, one per GPU, NCCL for the
plumbing; rank 0 compares the exchanged image with a single-GPU render bit for bit — for the NCCL gather path and for the
fused peer-store path (tools/multigpu_check.py is the worker)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("size", ["251x123", "640x360"])
def test_two_gpu_exchange_matches_single_gpu(size):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, SIZE=size)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "multigpu_check.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTIGPU_OK" in res.stdout, res.stdout[-3000:]
    assert res.stdout.count("exchanged image == single-GPU render: True") == 4     # NCCL gather, fused RGBA, fused RGB batched, the same with rotating roots
