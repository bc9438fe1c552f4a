#!/bin/bash
# round 2, 8-GPU call: sharded parity on 8 real peers, then the scaling family at N = 8, 4, 2 (34 / 33 / 32 qubits)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n8_gpus.txt
B2SV_TEST_WORLDS=8 timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q -p no:cacheprovider -k "sharded_gpu_matches_oracle" > gpurun_out/r2n8_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2n8_pytest.log
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 300 gpurun_out/r2_bench_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N ms/step',d.get('ms_per_step'),'weak_eff',d.get('weak_efficiency'),'dense',d.get('inputs',{}).get('dense'),'probe',d.get('exchange',{}).get('probe'),'check',d.get('check',{}).get('ok'),d.get('check',{}).get('parity_max_abs'), 'kern', d.get('kernels_rank0'))
except Exception as e: print('parse failed',e)
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 8 --steps 2 --warmup 1 --impl reference > gpurun_out/r2_bench_n8_reference.json 2> gpurun_out/r2_bench_n8_reference.err; echo "ref n8 rc=$?"; head -c 900 gpurun_out/r2_bench_n8_reference.json
