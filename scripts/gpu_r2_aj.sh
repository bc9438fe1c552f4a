#!/bin/bash
# round 2: more zero-fill store shapes (cfg 2 headline): memset, 128-bit streaming, 256-bit, plain 128-bit, 256-bit exact grid
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-extra-configs --no-dense"
for v in memset 0 2 3 4 memset 2 4; do
  unset B2SV_MEMSET B2SV_FILL
  if [ $v = memset ]; then export B2SV_MEMSET=1; else export B2SV_FILL=$v; fi
  $B > gpurun_out/r2aj_cfg2_$v.json 2>gpurun_out/r2aj_cfg2_$v.err || tail -n 3 gpurun_out/r2aj_cfg2_$v.err
  python - <<P
import json
d=json.load(open("gpurun_out/r2aj_cfg2_$v.json")); print("$v cfg2", round(d["ms_per_step"],4), {k:round(x["ms_per_step"],4) for k,x in d["kernels"].items()}, d["check"]["norm_last"])
P
done
