"""GPU, world_size 2 over NCCL (skipped on a single-GPU box): train_step.FlatGradSync and TrainStep(ddp=True), eager and captured in
a CUDA graph, through tools/check_flat_sync.py under `torch.distributed.run` -- rank-averaged gradients equal a single-process
backward (1e-5), identical weights on every rank after three steps.  The CPU twin is tests/test_ddp_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_flat_grad_sync():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_flat_sync.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "FlatGradSync gradients vs single-process backward" in r.stdout
    assert r.stdout.count("TrainStep graph=True") == 2 and r.stdout.count("TrainStep graph=False") == 2
