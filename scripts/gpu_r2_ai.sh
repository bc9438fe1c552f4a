#!/bin/bash
# round 2: zero-fill variants in the window-mode permutation sweep (cfg 2 headline): cudaMemsetAsync, own streaming stores, own TMA bulk stores
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-extra-configs --no-dense"
for v in memset stores bulk memset stores bulk; do
  unset B2SV_MEMSET B2SV_FILL
  if [ $v = memset ]; then export B2SV_MEMSET=1; fi
  if [ $v = bulk ]; then export B2SV_FILL=1; fi
  $B > gpurun_out/r2ai_cfg2_$v.json 2>gpurun_out/r2ai_cfg2_$v.err || tail -n 3 gpurun_out/r2ai_cfg2_$v.err
  python - <<P
import json
d=json.load(open("gpurun_out/r2ai_cfg2_$v.json")); print("$v cfg2", round(d["ms_per_step"],4), {k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()}, d["check"])
P
done
