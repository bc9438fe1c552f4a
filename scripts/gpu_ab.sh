#!/bin/bash
# A/B runs: bash scripts/gpu_ab.sh "<workload args>" "<variant args 1>" "<variant args 2>" ...
cd "$(dirname "$0")/.."
w=$1; shift
for v in "$@"; do
  echo "== $w | $v"
  timeout 900 python bench.py --workload $w --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline --no-e2e $v 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('gates/s %.0f  ms/step %.3f' % (d['value'], d['ms_per_step']))
for k,v in d['kernels'].items(): print('   %-9s launches/step %7.1f  ms/step %10.3f  frac %s' % (k, v['launches_per_step'], v['ms_per_step'], ('%.3f' % v['frac_of_peak']) if v['frac_of_peak'] else '-'))
"
done
