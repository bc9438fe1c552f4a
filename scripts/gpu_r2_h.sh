#!/bin/bash
# round 2, 1-GPU call: A/B of the per-SM compute gate (B2SV_GATE_LIMIT) on the heavy sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for L in 0 1 2; do
  for M in 3; do
    B2SV_GATE_LIMIT=$L B2SV_GATE_MIN_OPS=$M B2SV_DUMP_TIMES=gpurun_out/r2h_times_bpx_L$L.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2h_bpx30_L$L.log 2>&1; echo "bpx30 gate=$L rc=$?"; tail -n 1 gpurun_out/r2h_bpx30_L$L.log
    B2SV_GATE_LIMIT=$L B2SV_GATE_MIN_OPS=$M $P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --dense --jit sync > gpurun_out/r2h_cfg3_L$L.log 2>&1; echo "cfg3 gate=$L rc=$?"; tail -n 1 gpurun_out/r2h_cfg3_L$L.log
    B2SV_GATE_LIMIT=$L B2SV_GATE_MIN_OPS=$M $P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2h_pinv_L$L.log 2>&1; echo "pinv gate=$L rc=$?"; tail -n 1 gpurun_out/r2h_pinv_L$L.log
  done
done
B2SV_GATE_LIMIT=1 B2SV_GATE_MIN_OPS=1 B2SV_DUMP_TIMES=gpurun_out/r2h_times_bpx_L1M1.txt $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2h_bpx30_L1M1.log 2>&1; echo "bpx30 gate=1 min=1 rc=$?"; tail -n 1 gpurun_out/r2h_bpx30_L1M1.log
B2SV_GATE_LIMIT=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "goldens_complex128 or random" > gpurun_out/r2h_pytest_gate2.log 2>&1; echo "pytest gate=2 rc=$?"; tail -n 2 gpurun_out/r2h_pytest_gate2.log
