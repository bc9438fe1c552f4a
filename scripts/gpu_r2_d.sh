#!/bin/bash
# round 2, 1-GPU call: parity subset after the SWAP-frame planner change, BPX / pinv / cfg3 timings, ncu of a pure-BLK2
# sweep and a pure register-phase sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2d_pytest.log
P="python scripts/prof_apply.py"
B2SV_DUMP_TIMES=gpurun_out/r2d_times_bpx.txt $P cfg4_bpx_L8 --reps 2 > gpurun_out/r2d_bpx.log 2>&1; tail -n 1 gpurun_out/r2d_bpx.log
B2SV_AB=16 $P cfg4_bpx_L8 --reps 2 > gpurun_out/r2d_bpx_noframe.log 2>&1; tail -n 1 gpurun_out/r2d_bpx_noframe.log
$P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2d_bpx30.log 2>&1; tail -n 1 gpurun_out/r2d_bpx30.log
$P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2d_pinv.log 2>&1; tail -n 1 gpurun_out/r2d_pinv.log
$P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --jit sync > gpurun_out/r2d_cfg3.log 2>&1; tail -n 1 gpurun_out/r2d_cfg3.log
$P cfg3_gaussian_s3_t21 --dtype c64 --reps 2 --jit sync --dense > gpurun_out/r2d_cfg3_dense.log 2>&1; tail -n 1 gpurun_out/r2d_cfg3_dense.log
$P cfg4_bpx_L8 > gpurun_out/r2d_plain.log 2>&1 || { echo "plain failed"; exit 1; }
for L in 119 123; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_swz -s $L -c 1 -o gpurun_out/prof_r2d_bpx_$L -f $P cfg4_bpx_L8 > gpurun_out/r2d_ncu_$L.log 2>&1
  echo "ncu launch $L rc=$?"
done
du -sh gpurun_out
