#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "monomial" 2>&1 | tail -n 2
timeout 300 python scripts/microbench_mono.py 2>&1 | tee gpurun_out/microbench_mono.txt
rm -f gpurun_out/times_bpx.txt
B2SV_DUMP_TIMES=gpurun_out/times_bpx.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --workload cfg4_bpx_L8 | tail -n 1 | cut -c1-300
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload cfg2_laplace_n26 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})"
