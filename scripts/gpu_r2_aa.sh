#!/bin/bash
# round 2, 1-GPU call: bank-conflict-free lane assignment (lane_swaps) -- A/B timings, suite, conflict counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
for V in default B2SV_NO_LANE_SWAPS; do
  if [ "$V" = default ]; then E=""; else E="$V=1"; fi
  env $E $P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2aa_bpx30_$V.log 2>&1; echo "bpx30 $V rc=$?"; tail -n 1 gpurun_out/r2aa_bpx30_$V.log
  env $E $P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2aa_pinv_$V.log 2>&1; echo "pinv $V rc=$?"; tail -n 1 gpurun_out/r2aa_pinv_$V.log
  env $E $P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --dense --jit sync > gpurun_out/r2aa_cfg3d_$V.log 2>&1; echo "cfg3 dense $V rc=$?"; tail -n 1 gpurun_out/r2aa_cfg3d_$V.log
  env $E $P cfg2_laplace_n26 --reps 3 --jit sync --dense > gpurun_out/r2aa_cfg2d_$V.log 2>&1; echo "cfg2 dense $V rc=$?"; tail -n 1 gpurun_out/r2aa_cfg2d_$V.log
done
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2aa_pytest.log
for V in default B2SV_NO_LANE_SWAPS; do
  if [ "$V" = default ]; then E=""; else E="$V=1"; fi
  env $E timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:k_tile_swz -s 125 -c 2 --csv --log-file gpurun_out/r2aa_conflicts_$V.csv python scripts/prof_apply.py cfg4_bpx_L8 --qubits 30 > /dev/null 2>&1; echo "ncu conflicts $V rc=$?"
done
