#!/bin/bash
# round 2: cost of a light pass in the monomial-phase decision (measured 0.13 sweep, modelled 0.08 pass): A/B on BPX and the Pseudoinverse
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # tag base slope workload steps
  export B2SV_LIGHT_PASS=$2 B2SV_LIGHT_SLOPE=$3
  python bench.py --workload $4 --qubits 30 --steps $5 --warmup 1 --no-cpu-baseline --no-e2e --no-dense > gpurun_out/r2am_$4_$1.json 2> gpurun_out/r2am_$4_$1.err || tail -n 2 gpurun_out/r2am_$4_$1.err
  python - <<P
import json
d=json.load(open("gpurun_out/r2am_$4_$1.json")); print("$1 $4", round(d["ms_per_step"],3), {k:(round(x["ms_per_step"],2)) for k,x in d["kernels"].items()}, d["check"]["norm_last"])
P
}
run old 0.08 1.0 cfg4_bpx_L8 3
run A 0.7 1.0 cfg4_bpx_L8 3
run B 0.8 0.5 cfg4_bpx_L8 3
run old 0.08 1.0 cfg4_pinv_bpx_L4 1
run B 0.8 0.5 cfg4_pinv_bpx_L4 1
