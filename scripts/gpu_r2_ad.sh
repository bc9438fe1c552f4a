#!/bin/bash
# round 2: small multiplexors absorbed into dense blocks (cfg 3: 3 -> 2 passes per sweep).  A/B on cfg 3, whole GPU suite,
# the bench line, ncu --set full of the cfg 3 kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
C="python bench.py --workload cfg3_gaussian_s3_t21 --dtype c64 --qubits 28 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
B2SV_NO_MUX_IN_BLOCK=1 $C > gpurun_out/r2ad_cfg3_before.json 2> gpurun_out/r2ad_cfg3_before.err; echo "cfg3 before rc=$?"
$C > gpurun_out/r2ad_cfg3_after.json 2> gpurun_out/r2ad_cfg3_after.err; echo "cfg3 after rc=$?"
python - <<'P'
import json
for t in ("before", "after"):
    d = json.load(open(f"gpurun_out/r2ad_cfg3_{t}.json"))
    print(t, round(d["ms_per_step"], 4), {k: (round(v["ms_per_step"], 4), round(v["frac_of_peak"], 3)) for k, v in d["kernels"].items()}, "dense", round(d["inputs"]["dense"]["ms_per_step"], 4), {k: (round(v["ms_per_step"], 4), round(v["frac_of_peak"], 3)) for k, v in d["inputs"]["dense"]["kernels"].items()})
P
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2ad_pytest.log
timeout 900 python bench.py > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2ad_bench.err; cut -c1-200 gpurun_out/r2ad_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tile_swz|b2sv_perm_jit' -s 9 -c 3 -o gpurun_out/prof_r2c_cfg3 -f $C --no-dense > gpurun_out/r2ad_ncu_cfg3.log 2>&1; echo "ncu full cfg3 rc=$?"
