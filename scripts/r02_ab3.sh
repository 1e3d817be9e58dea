#!/bin/bash
# round 2 A/B call: smoke gate, parity subset on the default library, then kernel_ms of the default and of every build_variants/lib_*.so
mkdir -p gpurun_out
bl() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms'].items()})"; }
echo "== smoke gate"
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
echo "== parity (default library)"
timeout ${PT:-400} python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${PK:-implsch_matches or wamintgr_steps or depth_limited or golden or gravity_capillary or sea_ice or odd_nproma}" 2>&1 | tail -8
echo "== default library"
timeout 200 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-aux --no-extra 2>&1 | tail -1 | bl default
echo "== variants"
EXTRA="--no-aux --no-extra" bash scripts/bench_variants.sh
