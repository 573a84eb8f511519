"""Multi-GPU: sharded bundle adjustment over NCCL == single GPU (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_ba_equals_single_gpu(lib):
    n = lib.orbs_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(ROOT, "tests", "mgpu_ba_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "MGPU_BA_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
