#!/bin/bash
# round 2, 1-GPU call: member overhead 0 / 64 / 128 with the spill-free build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for C in 0 64 128; do
  B2SV_MEMBER_COST=$C $P cfg2_laplace_n26 --reps 3 --jit sync > gpurun_out/r2m_cfg2_C$C.log 2>&1; echo "cfg2 cost=$C rc=$?"; tail -n 1 gpurun_out/r2m_cfg2_C$C.log
  B2SV_MEMBER_COST=$C $P cfg2_laplace_n26 --reps 3 --jit sync --dense > gpurun_out/r2m_cfg2d_C$C.log 2>&1; echo "cfg2 dense cost=$C rc=$?"; tail -n 1 gpurun_out/r2m_cfg2d_C$C.log
  B2SV_MEMBER_COST=$C B2SV_DUMP_TIMES=gpurun_out/r2m_times_bpx30_C$C.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2m_bpx30_C$C.log 2>&1; echo "bpx30 cost=$C rc=$?"; tail -n 1 gpurun_out/r2m_bpx30_C$C.log
  B2SV_MEMBER_COST=$C $P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2m_pinv_C$C.log 2>&1; echo "pinv cost=$C rc=$?"; tail -n 1 gpurun_out/r2m_pinv_C$C.log
  B2SV_MEMBER_COST=$C $P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --dense --jit sync > gpurun_out/r2m_cfg3d_C$C.log 2>&1; echo "cfg3 dense cost=$C rc=$?"; tail -n 1 gpurun_out/r2m_cfg3d_C$C.log
done
