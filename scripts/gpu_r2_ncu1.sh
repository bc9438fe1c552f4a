#!/bin/bash
# round 2, call 1: ncu --set full of the heavy BPX tile sweeps (k_tile_swz<double2,1>) and of the cfg3 c64 kernels
# (gpurun_out/ must stay under 64 MiB: one launch per capture, raw pages exported on the box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt
P="python scripts/prof_apply.py"
B2SV_DUMP_TIMES=gpurun_out/r2_times_bpx.txt $P cfg4_bpx_L8 > gpurun_out/r2_plain_bpx.log 2>&1 || { echo "plain bpx failed"; tail -5 gpurun_out/r2_plain_bpx.log; exit 1; }
for L in 126 134 136; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_swz -s $L -c 1 -o gpurun_out/prof_r2_bpx_$L -f $P cfg4_bpx_L8 > gpurun_out/r2_ncu_bpx_$L.log 2>&1
  echo "ncu bpx launch $L rc=$?"
done
$P cfg3_gaussian_s3_t21 --dtype c64 --jit sync --dense > gpurun_out/r2_plain_cfg3.log 2>&1 || { echo "plain cfg3 failed"; tail -5 gpurun_out/r2_plain_cfg3.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_swz|b2sv_perm_jit|k_project_norm2' -c 5 -o gpurun_out/prof_r2_cfg3 -f $P cfg3_gaussian_s3_t21 --dtype c64 --jit sync --dense > gpurun_out/r2_ncu_cfg3.log 2>&1
echo "ncu cfg3 rc=$?"
tail -n 3 gpurun_out/r2_plain_bpx.log gpurun_out/r2_plain_cfg3.log
du -sh gpurun_out
