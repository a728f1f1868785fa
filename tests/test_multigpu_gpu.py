"""NCCL data-parallel parity (needs >= 2 GPUs on the box; skipped otherwise): launches tools/ddp_parity.py under
torchrun and checks the 2-rank product run against the r3d18_w2 fixture of the unmodified reference."""
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_matches_reference_fixture():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(ROOT / "tools" / "ddp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DDP PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
