#!/bin/bash
# round 2, 1-GPU call: sparse result hand-off (b2sv_project_gather_sparse), commuting-cluster pass, unscaled pass cost; suite + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2u_pytest.log
python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2u_bench.err
