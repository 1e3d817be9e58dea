"""bench.py's driver contract, the parts that run without a GPU: the reference arm prints ONE JSON line with the agreed keys
(the CPU restatement of the reference on a bounded sample), rank > 0 stays silent, and the B200 arm refuses to run without a
CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


@pytest.mark.parametrize("physics", ["default", "cy49r1"])
def test_reference_arm_prints_the_contract_line(built, physics):
    r = _run(["--impl", "reference", "--workload", "O48", "--steps", "1", "--warmup", "0", "--physics", physics])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "spectra/s" and d["value"] > 0
    assert d["metric"].startswith("grid-point spectra/s")
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent(built):
    r = _run(["--impl", "reference", "--workload", "O48", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_needs_a_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--workload", "O48", "--steps", "1", "--warmup", "3", "--no-cpu", "--no-e2e", "--no-aux"], timeout=300)
    assert r.returncode != 0
    assert "{\"metric\"" not in r.stdout                    # no number without the CUDA path
