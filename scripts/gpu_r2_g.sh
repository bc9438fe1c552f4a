#!/bin/bash
# round 2, 1-GPU call: range-checking debug build over the sanitizer cases (memcheck stand-in: compute-sanitizer is closed
# on this pool), then the ncu launch lists + one --set full capture that go with the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B2SV_LIB=$PWD/unitaria_b200/libunitaria_b200_dbg.so timeout 900 python scripts/sanitize_cases.py > gpurun_out/r2_bounds_check.log 2>&1; echo "bounds-check build rc=$?"; tail -n 2 gpurun_out/r2_bounds_check.log
B2SV_LIB=$PWD/unitaria_b200/libunitaria_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "monomial or live_window or goldens_complex128 and (fused or small_tiles) or embed or enumerate" > gpurun_out/r2_bounds_pytest.log 2>&1; echo "bounds-check pytest rc=$?"; tail -n 3 gpurun_out/r2_bounds_pytest.log
B="python bench.py --steps 3 --warmup 3 --no-extra-configs --no-cpu-baseline --no-e2e --no-dense"
$B > gpurun_out/r2g_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_cfg2.csv $B > gpurun_out/r2g_ncu_l1.log 2>&1; echo "ncu launches cfg2 rc=$?"
P="python scripts/prof_apply.py cfg4_bpx_L8 --qubits 30"
$P > gpurun_out/r2g_plain_bpx.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bpx30.csv $P > gpurun_out/r2g_ncu_l2.log 2>&1; echo "ncu launches bpx rc=$?"
$B > gpurun_out/r2g_plain_bench2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_swz|k_perm_scatter|k_project_norm2' -s 12 -c 4 -o gpurun_out/prof_r2_cfg2 -f $B > gpurun_out/r2g_ncu_f.log 2>&1; echo "ncu full cfg2 rc=$?"
du -sh gpurun_out
