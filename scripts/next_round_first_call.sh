#!/bin/bash
# What the first gpurun call of the next round should collect (everything that was written after this round's GPU budget ran out):
#   1. the late parity cases (cy50r1 combination, depth-limited points with the gravity-capillary physics)
#   2. A/B of the kernel experiments that are in the tree as compile-time switches (build them HERE first:
#        scripts/build_variants.sh base "" fifo "-DST_DP_FIFO=1"   ), default physics and cy49r1
# usage (on the GPU box):  bash scripts/next_round_first_call.sh > gpurun_out/first_call.log 2>&1
mkdir -p gpurun_out
echo "== late parity cases"
timeout 120 python -m pytest tests/test_gpu_zz_late_cases.py -q -m gpu 2>&1 | tail -5
if ls build_variants/lib_*.so >/dev/null 2>&1; then
  echo "== variants, default physics"
  bash scripts/bench_variants.sh
  echo "== variants, cy49r1"
  EXTRA="--physics cy49r1 --no-aux" bash scripts/bench_variants.sh
  echo "== parity of the fifo variant (same tests, other library)"
  [ -f build_variants/lib_fifo.so ] && ECWAM_B200_LIB=$PWD/build_variants/lib_fifo.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "implsch_matches or wamintgr_steps or golden or stencil_kernel" 2>&1 | tail -3
fi
