#!/bin/bash
# multi-GPU call: NCCL halo parity (device and host-buffer paths) and the bench line at N ranks
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 scripts/multirank_check.py 2>&1 | grep -v "^W\|^\[W" | tail -40 | tee gpurun_out/multirank_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu ${BX} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","kernel_ms")})
print("e2e",{k:v for k,v in d["e2e"].items() if k!="what"}); print("e2e_resident",{k:v for k,v in (d.get("e2e_resident") or {}).items() if k!="what"}); print("extra",d.get("extra"))
PY
