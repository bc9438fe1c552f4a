#!/bin/bash
# round 2, 1-GPU call: the driver's bench commands (own arm + reference arm), timed
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=$(date +%s); python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$? $(( $(date +%s) - S )) s"; tail -c 600 gpurun_out/r2s_bench.json
S=$(date +%s); python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err; echo "reference rc=$? $(( $(date +%s) - S )) s"; tail -c 400 gpurun_out/r2s_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2s_smoke.log
