#!/bin/bash
# round 2, final 1-GPU check: smoke, the whole GPU suite, the default bench line and the reference arm (as the driver runs them)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2ac_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2ac_pytest.log
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"; tail -n 3 gpurun_out/r2ac_bench.err; cut -c1-300 gpurun_out/r2ac_bench.json
S=$(date +%s); timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ac_bench_reference.json 2> gpurun_out/r2ac_ref.err; echo "reference arm rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/r2ac_bench_reference.json
