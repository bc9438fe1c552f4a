#!/bin/bash
# round 2: marginal cost of a shared-memory pass of each kind inside a compute-bound tile sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/microbench_passes.py --qubits 28 > gpurun_out/r2ae_passes.txt 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2ae_passes.txt | cut -c1-200
