#!/bin/bash
# round 2: own zero-fill kernel vs cudaMemsetAsync in the window-mode permutation sweep (cfg 2 headline, cfg 3), then parity subset
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-extra-configs --no-dense"
for v in memset own memset own; do
  if [ $v = memset ]; then export B2SV_MEMSET=1; else unset B2SV_MEMSET; fi
  $B > gpurun_out/r2ah_cfg2_$v.json 2>/dev/null
  python - <<P
import json
d=json.load(open("gpurun_out/r2ah_cfg2_$v.json")); print("$v cfg2", round(d["ms_per_step"],4), {k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()})
P
done
unset B2SV_MEMSET
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_distributed.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2ah_pytest.log
