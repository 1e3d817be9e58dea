#!/bin/bash
# Round profiling recipe (run under gpurun): bench line, ncu launch list, ncu --set full of the kernels.
# usage: scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json
# every launch of our kernels with its device time (cold-cache, serialised: compare SHARES) -- same command as the bench
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_point|k_stencil|propags2|copyback|pack_kernel|pad_kernel' -c 44 \
  --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-aux > gpurun_out/ncu_launch_${TAG}.log 2>&1
# full capture of the three kernels at the bench workload (O640): DRAM traffic per launch, pipe utilisation, stalls
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_point|k_stencil|propags2_kernel' -s 4 -c 4 -f -o gpurun_out/prof_${TAG}_O640 \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-aux > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
