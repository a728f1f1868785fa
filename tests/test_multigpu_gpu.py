"""NCCL data-parallel parity (needs >= W GPUs on the box; skipped otherwise): launches tools/ddp_parity.py under torchrun
and checks the W-rank product run against the r3d18_w{W} fixture of the unmodified reference — shuffled batches, gathered
keys and queue columns bit-exact, logits / loss / gradients within the stated bf16 tolerance.  Green logs of the same
command at W = 2 / 4 / 8 are committed under profiles/ (r02_ddp_parity_w*.txt)."""
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_ranks_match_reference_fixture(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run under gpurun --gpus {world})")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29533 + world), str(ROOT / "tools" / "ddp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"DDP PARITY OK (world {world})" in r.stdout, r.stdout[-4000:] + r.stderr[-3000:]
