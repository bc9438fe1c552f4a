#!/bin/bash
# round 2, call 2: the whole GPU suite (incl. two ranks on one GPU), BPX per-sweep times with / without the
# constant-bank block slots, and the N=1 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r2b_pytest.log
P="python scripts/prof_apply.py"
B2SV_DUMP_TIMES=gpurun_out/r2b_times_bpx.txt $P cfg4_bpx_L8 --reps 2 > gpurun_out/r2b_bpx.log 2>&1; tail -n 2 gpurun_out/r2b_bpx.log
B2SV_NO_BLK_SLOTS=1 B2SV_DUMP_TIMES=gpurun_out/r2b_times_bpx_noslots.txt $P cfg4_bpx_L8 --reps 2 > gpurun_out/r2b_bpx_noslots.log 2>&1; tail -n 2 gpurun_out/r2b_bpx_noslots.log
$P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2b_bpx30.log 2>&1; tail -n 2 gpurun_out/r2b_bpx30.log
$P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2b_pinv.log 2>&1; tail -n 2 gpurun_out/r2b_pinv.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2b_bench.err; head -c 3000 gpurun_out/r2b_bench.json
