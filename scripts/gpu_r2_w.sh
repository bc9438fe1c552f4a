#!/bin/bash
# round 2, 1-GPU call: A/B of the plain-stretch pass cost in the block-fusion DP (B2SV_PASS_COST)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for C in 0 24 48 96; do
  B2SV_PASS_COST=$C $P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2w_pinv_$C.log 2>&1; echo "pinv pass=$C rc=$?"; tail -n 1 gpurun_out/r2w_pinv_$C.log
  B2SV_PASS_COST=$C $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2w_bpx30_$C.log 2>&1; echo "bpx30 pass=$C rc=$?"; tail -n 1 gpurun_out/r2w_bpx30_$C.log
  B2SV_PASS_COST=$C $P cfg2_laplace_n26 --reps 3 --jit sync --dense > gpurun_out/r2w_cfg2d_$C.log 2>&1; echo "cfg2 dense pass=$C rc=$?"; tail -n 1 gpurun_out/r2w_cfg2d_$C.log
done
