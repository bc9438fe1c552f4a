#!/bin/bash
# round 2: last check of the committed build -- smoke + the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r2al_bench.json"))
print(d["ms_per_step"], d["value"], "launches", d["gpu_launches"], "e2e", d["e2e"]["value"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, d["clocks"])
P
