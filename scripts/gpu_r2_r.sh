#!/bin/bash
# round 2, 1-GPU call: commuting block merge (pairs of 4x4 blocks sharing a qubit -> one 8x8); suite + timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2r_pytest.log
P="timeout 300 python scripts/prof_apply.py"
B2SV_DUMP_TIMES=gpurun_out/r2r_times_bpx30.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2r_bpx30.log 2>&1; echo "bpx30 rc=$?"; tail -n 1 gpurun_out/r2r_bpx30.log
$P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2r_pinv.log 2>&1; echo "pinv rc=$?"; tail -n 1 gpurun_out/r2r_pinv.log
