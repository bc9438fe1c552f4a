#!/bin/bash
# round 2, 1-GPU call: the suite after the scatter form of window-mode permutation sweeps, then the N=1 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2e_pytest.log
timeout 900 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'] if d['e2e'] else None, d['e2e'].get('pipelined_columns') if d['e2e'] else None)
print('dense',{k:d['inputs']['dense'][k] for k in ('ms_per_step','gates_per_sec','frac')})
for k,v in d['configs'].items(): print(k,{kk:v.get(kk) for kk in ('ms_per_step','gates_per_sec','frac','error')}, {c:(round(x['ms_per_step'],3),round(x['frac_of_peak'] or 0,3)) for c,x in v.get('kernels',{}).items()})
print('kernels',{c:(round(x['ms_per_step'],3),round(x['frac_of_peak'] or 0,3)) for c,x in d['kernels'].items()})
PY
