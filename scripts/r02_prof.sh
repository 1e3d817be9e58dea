#!/bin/bash
# round 2: full ncu capture of the four hot kernels of the current build at the bench workload, then the reference arm on the same workload
mkdir -p gpurun_out
T=${TAG:-r02m}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'propags2|k_sweep|k_point' -s 12 -c 4 -f -o gpurun_out/prof_${T} \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aux --no-extra > gpurun_out/ncu_${T}.log 2>&1; tail -2 gpurun_out/ncu_${T}.log | cut -c1-300
if [ -n "$REF" ]; then
  free -g | head -2; nproc
  ( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_${T}.json 2> gpurun_out/bench_ref_${T}.err
  tail -c 1500 gpurun_out/bench_ref_${T}.json; tail -5 gpurun_out/bench_ref_${T}.err
fi
