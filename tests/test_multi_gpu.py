"""Runs tests/multi_gpu_check.py under torchrun when the box has at least two GPUs (skipped otherwise; the host-side
collective algebra is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_paths_match_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MULTI-GPU CHECK OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
