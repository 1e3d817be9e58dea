#!/bin/bash
# time every build_variants/lib_*.so on the GPU box: bench.py kernel_ms per variant (developer tool)
WL=${WL:-O640}
for f in build_variants/lib_*.so; do
  n=$(basename $f .so)
  ECWAM_B200_LIB=$PWD/$f timeout 200 python bench.py --workload $WL --steps 4 --warmup 3 --no-e2e --no-cpu $EXTRA 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['kernel_ms'].items()}, {k: round(v,2) for k,v in (d.get('output_step') or {}).items() if k.endswith('_ms')})" || echo "$n FAILED"
done
