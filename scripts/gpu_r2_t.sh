#!/bin/bash
# round 2, 1-GPU call: suite, measured c64 errors, bench line, ncu launch lists + --set full captures for profiles/, Pseudoinverse per-launch times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2t_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -p no:cacheprovider -k columns 2>&1 | grep "max relative" > gpurun_out/r2t_fullsize_errors.txt; cat gpurun_out/r2t_fullsize_errors.txt
python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-extra-configs --no-cpu-baseline --no-e2e --no-dense"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b_cfg2.csv $B > gpurun_out/r2t_ncu_l1.log 2>&1; echo "ncu launches cfg2 rc=$?"
P="python scripts/prof_apply.py cfg4_bpx_L8 --qubits 30"
B2SV_DUMP_TIMES=gpurun_out/r2t_times_bpx30.txt $P > gpurun_out/r2t_plain_bpx.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b_bpx30.csv $P > gpurun_out/r2t_ncu_l2.log 2>&1; echo "ncu launches bpx rc=$?"
for L in 125 126; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_swz -s $L -c 1 -o gpurun_out/prof_r2b_bpx30_$L -f $P > gpurun_out/r2t_ncu_bpx_$L.log 2>&1; echo "ncu full bpx launch $L rc=$?"
done
C="python scripts/prof_apply.py cfg3_gaussian_s3_t21 --dtype c64 --jit sync --dense"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_swz|b2sv_perm_jit' -c 3 -o gpurun_out/prof_r2b_cfg3 -f $C > gpurun_out/r2t_ncu_cfg3.log 2>&1; echo "ncu full cfg3 rc=$?"
B2SV_DUMP_TIMES=gpurun_out/r2t_times_pinv30.txt timeout 300 python scripts/prof_apply.py cfg4_pinv_bpx_L4 --qubits 30 --jit sync > gpurun_out/r2t_pinv30.log 2>&1; echo "pinv30 rc=$?"; tail -n 1 gpurun_out/r2t_pinv30.log
du -sh gpurun_out
