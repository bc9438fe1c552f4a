#!/bin/bash
# round 2, 1-GPU call: branch-free norm scan kernel -- microbench, suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/microbench_project.py 30 > gpurun_out/r2z_microbench_project.txt 2>&1; echo "microbench rc=$?"; cat gpurun_out/r2z_microbench_project.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2z_pytest.log
