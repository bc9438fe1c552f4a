#!/bin/bash
# round 2, 1-GPU call: blocks stored real with their common phase factored out and merged; suite + timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2q_pytest.log
P="timeout 300 python scripts/prof_apply.py"
B2SV_DUMP_TIMES=gpurun_out/r2q_times_bpx30.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2q_bpx30.log 2>&1; echo "bpx30 rc=$?"; tail -n 1 gpurun_out/r2q_bpx30.log
$P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2q_pinv.log 2>&1; echo "pinv rc=$?"; tail -n 1 gpurun_out/r2q_pinv.log
