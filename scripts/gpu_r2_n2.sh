#!/bin/bash
# round 2, 2-GPU call: the N=2 bench line with the final build (sanity of the sharded path after the planner changes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2af_bench_n2.json 2> gpurun_out/r2af_bench_n2.err; echo "bench n2 rc=$?"; tail -c 300 gpurun_out/r2af_bench_n2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2af_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 ms/step',d.get('ms_per_step'),'weak_eff',d.get('weak_efficiency'),'dense',d.get('inputs',{}).get('dense'),'check',d.get('check',{}).get('ok'),d.get('check',{}).get('parity_max_abs'), 'kern', d.get('kernels_rank0'))
except Exception as e: print('parse failed',e)
PY
