#!/bin/bash
# round 2, 8-GPU call: the N=8 and N=4 bench lines with the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for N in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2x_bench_n$N.json 2> gpurun_out/r2x_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 200 gpurun_out/r2x_bench_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2x_bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N ms/step',d.get('ms_per_step'),'weak_eff',d.get('weak_efficiency'),'dense',d.get('inputs',{}).get('dense',{}).get('ms_per_step'),'check',d.get('check',{}).get('ok'),d.get('check',{}).get('parity_max_abs'), 'kern', d.get('kernels_rank0'))
except Exception as e: print('parse failed',e)
PY
done
