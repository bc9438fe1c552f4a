#!/bin/bash
# round 2, final build (own zero-fill kernel): smoke, whole GPU suite, default bench line, ncu launch list of the bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ak_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2ak_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2ak_pytest.log
timeout 900 python bench.py > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/r2ak_bench.err; cut -c1-200 gpurun_out/r2ak_bench.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra-configs --no-dense"
$B > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2d_cfg2.csv $B > gpurun_out/r2ak_ncu_l1.log 2>&1; echo "ncu launches cfg2 rc=$?"
