#!/bin/bash
# round 2, call with the warp-specialised sweep: smoke gate, parity subset, A/B benches, ncu captures (every step under a timeout)
mkdir -p gpurun_out
T=${TAG:-r02d}
bl() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms'].items()})"; }
echo "== smoke gate"
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
echo "== parity (default library)"
timeout ${PT:-300} python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${PK:-implsch_matches or wamintgr_steps or stencil_kernel or depth_limited or golden or propags2}" 2>&1 | tail -15
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "PARITY FAILED"; fi
echo "== default library"
timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-aux 2>&1 | tail -1 | bl default
echo "== one-role sweep + exact propags2"
ECWAM_B200_PROPAG=exact ECWAM_B200_STENCIL=sweep timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-aux 2>&1 | tail -1 | bl sweep_exact
echo "== variants"
EXTRA="--no-aux" bash scripts/bench_variants.sh
if [ -n "$NCU" ]; then
echo "== ncu"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'propags2|k_sweep|k_point' -s 12 -c 4 -f -o gpurun_out/prof_${T} \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aux > gpurun_out/ncu_${T}.log 2>&1; tail -2 gpurun_out/ncu_${T}.log
ECWAM_B200_PROPAG=exact timeout 300 ncu --set full --clock-control none --import-source on -k regex:'propags2' -s 3 -c 1 -f -o gpurun_out/prof_${T}_pexact \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aux > gpurun_out/ncu_${T}_pexact.log 2>&1; tail -2 gpurun_out/ncu_${T}_pexact.log
fi
echo "== default lib, old fast propags2 (fast1)"
ECWAM_B200_PROPAG=fast1 timeout 200 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-aux 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fast1', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms'].items()})"
