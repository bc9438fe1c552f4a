#!/bin/bash
# round 2, 1-GPU call: complex64 tile sweeps in fp32 arithmetic -- parity (goldens + full-size columns) and timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2j_pytest.log
P="timeout 300 python scripts/prof_apply.py"
$P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --dense --jit sync > gpurun_out/r2j_cfg3_dense.log 2>&1; echo "cfg3 dense rc=$?"; tail -n 1 gpurun_out/r2j_cfg3_dense.log
$P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --jit sync > gpurun_out/r2j_cfg3.log 2>&1; echo "cfg3 rc=$?"; tail -n 1 gpurun_out/r2j_cfg3.log
$P cfg4_bpx_L8 --dtype c64 --reps 2 --qubits 30 > gpurun_out/r2j_bpx30_c64.log 2>&1; echo "bpx30 c64 rc=$?"; tail -n 1 gpurun_out/r2j_bpx30_c64.log
$P cfg4_pinv_bpx_L4 --dtype c64 --reps 2 --jit sync > gpurun_out/r2j_pinv_c64.log 2>&1; echo "pinv c64 rc=$?"; tail -n 1 gpurun_out/r2j_pinv_c64.log
