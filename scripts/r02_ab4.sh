#!/bin/bash
# A/B of an env-switched variant on one GPU: smoke gate, parity subset, then kernel_ms for O640 and O320 with the switch on / off
bl() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms'].items()})"; }
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${PK:-implsch_matches or wamintgr_steps or depth_limited or golden or gravity_capillary or single_chunk or host_buffer}" 2>&1 | tail -3
for wl in O640 O320; do
  for v in ${VALS:-1 0}; do
    env ${VAR}=$v timeout 200 python bench.py --workload $wl --steps 6 --warmup 3 --no-e2e --no-cpu --no-aux --no-extra 2>&1 | tail -1 | bl "$wl ${VAR}=$v"
  done
done
