#!/bin/bash
# round 2, 1-GPU call: first-wave stagger experiment (B2SV_STAGGER_US)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for S in 0 2 4 8; do
  B2SV_STAGGER_US=$S B2SV_DUMP_TIMES=gpurun_out/r2o_times_bpx30_S$S.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2o_bpx30_S$S.log 2>&1; echo "bpx30 stagger=$S rc=$?"; tail -n 1 gpurun_out/r2o_bpx30_S$S.log
done
