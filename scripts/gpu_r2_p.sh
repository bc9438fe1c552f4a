#!/bin/bash
# round 2, 1-GPU call: phase traces (load / compute / store timestamps per CTA) of chosen BPX sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for L in $TRACE_LAUNCHES; do
  B2SV_TRACE=gpurun_out/r2p_trace_$L.bin B2SV_TRACE_LAUNCH=$L $P cfg4_bpx_L8 --reps 1 --qubits 30 > gpurun_out/r2p_bpx30_$L.log 2>&1; echo "trace launch $L rc=$?"; tail -n 1 gpurun_out/r2p_bpx30_$L.log
done
ls -la gpurun_out/r2p_trace_*
