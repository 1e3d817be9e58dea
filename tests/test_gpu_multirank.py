"""Multi-GPU parity (pytest -m gpu, needs >= 2 GPUs; skipped on a 1-GPU box): N ranks with the NCCL halo exchange give the
same spectra as the 1-rank oracle.  Launched as torchrun-style subprocesses, one per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_halo_matches_oracle(built, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), os.path.join(ROOT, "scripts", "multirank_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTIRANK OK" in out.stdout
