#!/bin/bash
# Short device-time benchmarks of every BASELINE config (no CPU baseline, no e2e): one line each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/bench_all.txt
for w in "cfg2_laplace_n26" "cfg2_laplace_readme_n26" "cfg3_gaussian_s3_t21 --dtype c64" "cfg4_bpx_L8" "cfg4_bpx_L8 --qubits 30" "cfg4_pinv_bpx_L4" ${EXTRA_WORKLOADS}; do
  echo "== $w" | tee -a gpurun_out/bench_all.txt
  timeout 900 python bench.py --workload $w --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('gates/s %.0f  ms/step %.3f  live gates %d' % (d['value'], d['ms_per_step'], d['gates_live_per_step']))
for k,v in d['kernels'].items(): print('   %-9s launches/step %7.1f  ms/step %10.3f  frac %s' % (k, v['launches_per_step'], v['ms_per_step'], ('%.3f' % v['frac_of_peak']) if v['frac_of_peak'] else '-'))
" | tee -a gpurun_out/bench_all.txt
done
