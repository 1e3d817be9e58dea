for mode in q exact fast1; do
  ECWAM_B200_PROPAG=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu --no-e2e --no-aux --no-extra 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mode', round(d['ms_per_step'],2), {k: round(v,3) for k,v in d['kernel_ms'].items()})"
done
