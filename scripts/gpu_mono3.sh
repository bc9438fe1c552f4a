#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
summ='
import json,sys
d=json.loads(sys.stdin.read())
print("gates/s %.0f  ms/step %.3f" % (d["value"], d["ms_per_step"]), {k:(v["launches_per_step"], round(v["ms_per_step"],2)) for k,v in d["kernels"].items()})'
B="timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload"
: > gpurun_out/mono3.txt
for cost in 0.008 0.015 0.03; do
  echo "== pinv cost=$cost" | tee -a gpurun_out/mono3.txt
  B2SV_MONO_COST=$cost $B cfg4_pinv_bpx_L4 | tail -n 1 | python -c "$summ" | tee -a gpurun_out/mono3.txt
done
echo "== pinv no-mono" | tee -a gpurun_out/mono3.txt
B2SV_AB=8 $B cfg4_pinv_bpx_L4 | tail -n 1 | python -c "$summ" | tee -a gpurun_out/mono3.txt
rm -f gpurun_out/times_pinv.txt
B2SV_DUMP_TIMES=gpurun_out/times_pinv.txt timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --workload cfg4_pinv_bpx_L4 > /dev/null
for cost in 0.008 0.015 0.03; do
  echo "== bpx cost=$cost" | tee -a gpurun_out/mono3.txt
  B2SV_MONO_COST=$cost $B cfg4_bpx_L8 | tail -n 1 | python -c "$summ" | tee -a gpurun_out/mono3.txt
done
