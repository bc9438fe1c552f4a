#!/bin/bash
# round 2, 1-GPU call: A/B of the per-member overhead in the block-fusion cost model (B2SV_MEMBER_COST)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for C in 0 8 16 32 64; do
  B2SV_MEMBER_COST=$C B2SV_DUMP_TIMES=gpurun_out/r2k_times_bpx30_C$C.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2k_bpx30_C$C.log 2>&1; echo "bpx30 cost=$C rc=$?"; tail -n 1 gpurun_out/r2k_bpx30_C$C.log
  B2SV_MEMBER_COST=$C $P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2k_pinv_C$C.log 2>&1; echo "pinv cost=$C rc=$?"; tail -n 1 gpurun_out/r2k_pinv_C$C.log
done
for C in 16 32; do
  B2SV_MEMBER_COST=$C timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "goldens_complex128 and fused" > gpurun_out/r2k_pytest_C$C.log 2>&1; echo "pytest cost=$C rc=$?"; tail -n 2 gpurun_out/r2k_pytest_C$C.log
done
