"""Launches tests/multigpu_check.py under torchrun when the box has >= 2 GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_gpu_sharding_stats_and_pooled_adaptation():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
