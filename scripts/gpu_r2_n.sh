#!/bin/bash
# round 2, 1-GPU call: swizzle-once addressing; suite + timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2n_pytest.log
P="timeout 300 python scripts/prof_apply.py"
B2SV_DUMP_TIMES=gpurun_out/r2n_times_bpx30.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2n_bpx30.log 2>&1; echo "bpx30 rc=$?"; tail -n 1 gpurun_out/r2n_bpx30.log
$P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2n_pinv.log 2>&1; echo "pinv rc=$?"; tail -n 1 gpurun_out/r2n_pinv.log
$P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --dense --jit sync > gpurun_out/r2n_cfg3d.log 2>&1; echo "cfg3 dense rc=$?"; tail -n 1 gpurun_out/r2n_cfg3d.log
$P cfg2_laplace_n26 --reps 3 --jit sync --dense > gpurun_out/r2n_cfg2d.log 2>&1; echo "cfg2 dense rc=$?"; tail -n 1 gpurun_out/r2n_cfg2d.log
