#!/bin/bash
# full validation call: smoke, the whole GPU suite, the default bench line (with e2e, CPU sample and the extra configurations)
mkdir -p gpurun_out
T=${TAG:-r02h}
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "SMOKE FAILED: stopping"; exit 1; fi
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_${T}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 400 gpurun_out/bench_${T}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${T}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","kernel_ms")})
print("e2e",d["e2e"]); print("extra",d.get("extra")); print("cpu",d.get("cpu_baseline"))
print("roofline",{k:v for k,v in d["roofline"].items() if k not in ("note","per_kernel")})
for r in d["roofline"]["per_kernel"]: print(r)
PY
