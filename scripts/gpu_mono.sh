#!/bin/bash
# Round-1 late check of the monomial-phase tile ops: parity first, then A/B device-time benchmarks.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/mono_summary.txt
run() { # name, timeout, command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/mono_summary.txt
  timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/mono_summary.txt
  return $rc
}
run m_mono 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "monomial" || { tail -n 40 gpurun_out/m_mono.log; exit 1; }
run m_parity 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu -x || { tail -n 40 gpurun_out/m_parity.log; }
run m_smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
bench() { # label, env..., -- workload args
  local label=$1; shift
  echo "== $label" | tee -a gpurun_out/mono_bench.txt
  env "$@" | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('gates/s %.0f  ms/step %.3f  live gates %d' % (d['value'], d['ms_per_step'], d['gates_live_per_step']))
for k,v in d['kernels'].items(): print('   %-9s launches/step %7.1f  ms/step %10.3f  frac %s' % (k, v['launches_per_step'], v['ms_per_step'], ('%.3f' % v['frac_of_peak']) if v['frac_of_peak'] else '-'))
" 2>&1 | tee -a gpurun_out/mono_bench.txt
}
: > gpurun_out/mono_bench.txt
B="timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload"
bench "bpx_L8 mono"            B2SV_AB=0 $B cfg4_bpx_L8
bench "bpx_L8 no-mono"         B2SV_AB=8 $B cfg4_bpx_L8
bench "pinv mono cost=0.008"   B2SV_AB=0 $B cfg4_pinv_bpx_L4
bench "pinv no-mono"           B2SV_AB=8 $B cfg4_pinv_bpx_L4
bench "pinv mono cost=0.003"   B2SV_MONO_COST=0.003 $B cfg4_pinv_bpx_L4
bench "pinv mono cost=0.02"    B2SV_MONO_COST=0.02 $B cfg4_pinv_bpx_L4
bench "bpx_L8 30q mono"        B2SV_AB=0 $B cfg4_bpx_L8 --qubits 30
bench "cfg5_bits24"            B2SV_AB=0 $B cfg5_bits24
bench "cfg2"                   B2SV_AB=0 $B cfg2_laplace_n26
bench "cfg3 c64"               B2SV_AB=0 $B cfg3_gaussian_s3_t21 --dtype c64
cat gpurun_out/mono_summary.txt
