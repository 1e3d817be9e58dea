#!/bin/bash
# one gpurun call of round 2: smoke gate, parity of the default kernels, peaks, A/B of the kernel variants in build_variants/
# (every step under its own timeout: a hung kernel must not eat the call)
mkdir -p gpurun_out
echo "== host"; nproc; free -g | head -2; lscpu | grep "Model name"
echo "== smoke gate"
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
echo "== peaks"
timeout 120 python - <<'PY'
import ctypes as C
from ecwam_b200 import lib as L
lib = L.load()
a, b = C.c_double(), C.c_double()
print("measure_peaks rc", lib.ecwam_b200_measure_peaks(C.byref(a), C.byref(b)), "fp64 TFLOP/s", a.value, "copy GB/s", b.value)
PY
echo "== parity (default library)"
timeout ${PT:-400} python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${PK:-implsch_matches or wamintgr_steps or stencil_kernel or depth_limited or golden or propags2 or substeps or depth_refraction}" 2>&1 | tail -15
if [ -n "$FULLTESTS" ]; then timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8; fi
echo "== default library, dp kernel + exact propags2"
ECWAM_B200_PROPAG=exact ECWAM_B200_STENCIL=dp timeout 200 python bench.py --workload ${WL:-O640} --steps 4 --warmup 3 --no-e2e --no-cpu --no-aux 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('dp', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms'].items()})"
echo "== default library"
timeout 200 python bench.py --workload ${WL:-O640} --steps 4 --warmup 3 --no-e2e --no-cpu --no-aux 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms'].items()})"
echo "== variants"
EXTRA="--no-aux" bash scripts/bench_variants.sh
