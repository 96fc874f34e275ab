"""Two ranks over NCCL on one box (skipped with fewer than 2 GPUs): tools/dist_check.py — the sharded GRBCM's prediction,
all-gathered weights and GRBCM_MLE NLML/gradient equal the unsharded model's on every rank.
Last run on 2 x B200: profiles/r01_dist_check_2gpu.log (worst relative difference 6e-16)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_unsharded_over_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0 and "dist_check: ok" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
