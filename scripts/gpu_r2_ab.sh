#!/bin/bash
# round 2, 1-GPU call: lane assignment for monomial phases -- timings, conflict counters, suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P="timeout 300 python scripts/prof_apply.py"
$P cfg4_bpx_L8 --reps 2 --qubits 30 > gpurun_out/r2ab_bpx30.log 2>&1; echo "bpx30 rc=$?"; tail -n 1 gpurun_out/r2ab_bpx30.log
$P cfg4_pinv_bpx_L4 --reps 2 --jit sync > gpurun_out/r2ab_pinv.log 2>&1; echo "pinv rc=$?"; tail -n 1 gpurun_out/r2ab_pinv.log
$P cfg3_gaussian_s3_t21 --dtype c64 --reps 3 --dense --jit sync > gpurun_out/r2ab_cfg3d.log 2>&1; echo "cfg3 dense rc=$?"; tail -n 1 gpurun_out/r2ab_cfg3d.log
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:k_tile_swz -s 125 -c 2 --csv --log-file gpurun_out/r2ab_conflicts.csv python scripts/prof_apply.py cfg4_bpx_L8 --qubits 30 > /dev/null 2>&1; echo "ncu conflicts rc=$?"; grep -v "^==" gpurun_out/r2ab_conflicts.csv | cut -d, -f1,13-15 | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2ab_pytest.log
