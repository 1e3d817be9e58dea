"""Multi-rank parity (pytest -m gpu): N ranks give the same spectra as the 1-rank oracle.
  * NCCL halo exchange, one GPU per rank: needs >= N GPUs (skipped on a 1-GPU box);
  * the halo exchange supplied by the host (ecwam_b200_set_exchange, staged over gloo): the ranks SHARE the GPUs that are there,
    so the MPDECOMP tables, the pack kernel, the in-place receive layout and PROENVHALO are exercised on a 1-GPU box as well.
Launched as torchrun-style subprocesses."""
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


@pytest.mark.parametrize("world", [2, 3])
def test_host_supplied_halo_exchange_matches_oracle(built, world):
    """N processes on whatever GPUs the box has (all on cuda:0 on the driver's 1-GPU box), halo through the exchange callback."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29720 + world), os.path.join(ROOT, "scripts", "multirank_staged_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTIRANK STAGED OK" in out.stdout
