#!/bin/bash
# round 2, final build: ncu launch lists (gpu__time_duration) of the bench command for cfg 2 and cfg 3, each after the plain run exited 0
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-dense"
$B > gpurun_out/r2ag_plain_cfg2.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c_cfg2.csv $B > gpurun_out/r2ag_ncu_l1.log 2>&1; echo "ncu launches cfg2 rc=$?"
C="$B --workload cfg3_gaussian_s3_t21 --dtype c64 --qubits 28"
$C > gpurun_out/r2ag_plain_cfg3.json 2>/dev/null && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c_cfg3.csv $C > gpurun_out/r2ag_ncu_l2.log 2>&1; echo "ncu launches cfg3 rc=$?"
