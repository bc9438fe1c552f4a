#!/bin/bash
# Round-end evidence run on one B200: full bench line, reference arm, per-config device times, ncu launch
# lists and one ncu --set full capture of the tile kernel (each ncu pass only after the plain run exited 0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
T=${TAG:-r1_v6}
timeout 900 python bench.py > gpurun_out/bench_$T.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_$T.log > gpurun_out/bench_$T.json; cut -c1-400 gpurun_out/bench_$T.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 1 > gpurun_out/bench_${T}_reference.json; cut -c1-300 gpurun_out/bench_${T}_reference.json
EXTRA_WORKLOADS="cfg5_bits24" STEPS=3 bash scripts/gpu_bench_all.sh > /dev/null 2>&1; cp gpurun_out/bench_all.txt gpurun_out/bench_all_$T.txt; cat gpurun_out/bench_all_$T.txt
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}_cfg2.csv $B > gpurun_out/ncu_l1.log 2>&1; echo "ncu launches cfg2 rc=$?"
$B --workload cfg4_bpx_L8 > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${T}_bpx.csv $B --workload cfg4_bpx_L8 > gpurun_out/ncu_l2.log 2>&1; echo "ncu launches bpx rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tile_swz|b2sv_perm_jit|k_project_norm2' -s 15 -c 5 -o gpurun_out/prof_${T}_cfg2 -f $B > gpurun_out/ncu_f1.log 2>&1; echo "ncu full cfg2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tile_swz' -s 600 -c 8 -o gpurun_out/prof_${T}_bpx -f $B --workload cfg4_bpx_L8 > gpurun_out/ncu_f2.log 2>&1; echo "ncu full bpx rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -n 4
