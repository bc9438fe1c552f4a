#!/bin/bash
# round 2: parity of the final plans (light-pass cost from the microbenchmark) -- full-size columns + the parity suite, bounded
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 215 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_distributed.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2an_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2an_pytest.log
