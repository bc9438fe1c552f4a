#!/bin/bash
# round 2, 1-GPU call: once-per-op index deposit (tile_for_each), member overhead default 64; suite + timings + higher overheads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2l_pytest.log
P="timeout 300 python scripts/prof_apply.py"
for C in 64 128 256; do
  B2SV_MEMBER_COST=$C B2SV_DUMP_TIMES=gpurun_out/r2l_times_bpx30_C$C.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2l_bpx30_C$C.log 2>&1; echo "bpx30 cost=$C rc=$?"; tail -n 1 gpurun_out/r2l_bpx30_C$C.log
  B2SV_MEMBER_COST=$C $P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2l_pinv_C$C.log 2>&1; echo "pinv cost=$C rc=$?"; tail -n 1 gpurun_out/r2l_pinv_C$C.log
done
$P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --dense --jit sync > gpurun_out/r2l_cfg3_dense.log 2>&1; echo "cfg3 dense rc=$?"; tail -n 1 gpurun_out/r2l_cfg3_dense.log
$P cfg2_laplace_n26 --reps 3 --jit sync > gpurun_out/r2l_cfg2.log 2>&1; echo "cfg2 rc=$?"; tail -n 1 gpurun_out/r2l_cfg2.log
