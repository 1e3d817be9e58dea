#!/bin/bash
# round 2 evidence call: ncu launch list of the bench command, ncu --set full of the four hot kernels, host-band sweep of the e2e leg
mkdir -p gpurun_out
T=${TAG:-r02y}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_point|k_sweep|k_stencil|propags2|copyback|pack_kernel|pad_kernel|k_enh' -c 60 \
  --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-aux --no-extra > gpurun_out/ncu_launch_${T}.log 2>&1
tail -3 gpurun_out/launches_${T}.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'propags2|k_sweep|k_point' -s 12 -c 4 -f -o gpurun_out/prof_${T} \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-aux --no-extra > gpurun_out/ncu_${T}.log 2>&1; tail -2 gpurun_out/ncu_${T}.log | cut -c1-200
for nb in 24 48 96; do
  ECWAM_B200_HOST_BANDS=$nb timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-aux --no-extra 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bands $nb: e2e', round(d['e2e']['value']/1e6,3), 'M spectra/s; resident', round((d.get('e2e_resident') or {}).get('value',0)/1e6,2))"
done
