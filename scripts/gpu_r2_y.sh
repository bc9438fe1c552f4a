#!/bin/bash
# round 2, 1-GPU call: range-checking debug build over the sanitizer cases + parity subset (final kernels), then the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B2SV_LIB=$PWD/unitaria_b200/libunitaria_b200_dbg.so timeout 900 python scripts/sanitize_cases.py > gpurun_out/r2y_bounds_check.log 2>&1; echo "bounds-check cases rc=$?"; tail -n 2 gpurun_out/r2y_bounds_check.log
B2SV_LIB=$PWD/unitaria_b200/libunitaria_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "monomial or live_window or goldens_complex128 and (fused or small_tiles) or embed or enumerate or sparse" > gpurun_out/r2y_bounds_pytest.log 2>&1; echo "bounds-check pytest rc=$?"; tail -n 3 gpurun_out/r2y_bounds_pytest.log
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2y_bench.err
