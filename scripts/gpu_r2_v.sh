#!/bin/bash
# round 2, 1-GPU call: A/B of the commuting-cluster pass and the block merge on Pseudoinverse / BPX
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for V in default B2SV_NO_CLUSTER B2SV_NO_BLOCK_MERGE; do
  if [ "$V" = default ]; then E=""; else E="$V=1"; fi
  env $E $P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2v_pinv_$V.log 2>&1; echo "pinv $V rc=$?"; tail -n 1 gpurun_out/r2v_pinv_$V.log
  env $E $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2v_bpx30_$V.log 2>&1; echo "bpx30 $V rc=$?"; tail -n 1 gpurun_out/r2v_bpx30_$V.log
done
