#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "monomial or goldens_complex128 or mid_size" 2>&1 | tail -n 3
summ='
import json,sys
d=json.loads(sys.stdin.read())
print("gates/s %.0f  ms/step %.3f" % (d["value"], d["ms_per_step"]), {k:(v["launches_per_step"], round(v["ms_per_step"],2)) for k,v in d["kernels"].items()})'
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload"
: > gpurun_out/mono4.txt
run() { echo "== $1" | tee -a gpurun_out/mono4.txt; shift; env "$@" | tail -n 1 | python -c "$summ" | tee -a gpurun_out/mono4.txt; }
run "pinv tables" B2SV_AB=0 $B cfg4_pinv_bpx_L4
run "pinv chunk7" B2SV_CHUNK_BITS=7 $B cfg4_pinv_bpx_L4
run "pinv chunk5" B2SV_CHUNK_BITS=5 $B cfg4_pinv_bpx_L4
run "bpx tables" B2SV_AB=0 $B cfg4_bpx_L8
run "bpx chunk6" B2SV_CHUNK_BITS=6 $B cfg4_bpx_L8
run "bpx chunk3" B2SV_CHUNK_BITS=3 $B cfg4_bpx_L8
run "cfg2 chunk7" B2SV_CHUNK_BITS=7 $B cfg2_laplace_n26
run "cfg2" B2SV_AB=0 $B cfg2_laplace_n26
rm -f gpurun_out/times_pinv.txt gpurun_out/times_bpx.txt
B2SV_DUMP_TIMES=gpurun_out/times_pinv.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --workload cfg4_pinv_bpx_L4 > /dev/null
B2SV_DUMP_TIMES=gpurun_out/times_bpx.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --workload cfg4_bpx_L8 > /dev/null
timeout 300 python scripts/microbench_mono.py 2>&1 | grep -v "no-" | tee gpurun_out/microbench_mono_tab.txt
run "cfg3 c64" B2SV_AB=0 $B cfg3_gaussian_s3_t21 --dtype c64
run "cfg3 c64 chunk8" B2SV_CHUNK_BITS=8 $B cfg3_gaussian_s3_t21 --dtype c64
run "cfg5" B2SV_AB=0 $B cfg5_bits24
run "readme" B2SV_AB=0 $B cfg2_laplace_readme_n26
run "bpx30" B2SV_AB=0 $B cfg4_bpx_L8 --qubits 30
