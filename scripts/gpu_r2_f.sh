#!/bin/bash
# round 2, 1-GPU call: suite after the phase pre-decode / tail window / layout changes + BPX, pinv timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2f_pytest.log
P="python scripts/prof_apply.py"
B2SV_DUMP_TIMES=gpurun_out/r2f_times_bpx.txt $P cfg4_bpx_L8 --reps 2 > gpurun_out/r2f_bpx.log 2>&1; tail -n 1 gpurun_out/r2f_bpx.log
$P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2f_bpx30.log 2>&1; tail -n 1 gpurun_out/r2f_bpx30.log
$P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2f_pinv.log 2>&1; tail -n 1 gpurun_out/r2f_pinv.log
