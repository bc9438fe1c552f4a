#!/bin/bash
# Staged GPU check: simplest kernels first, so a fault in a later stage is easy to place.
# Usage (on the GPU box, via gpurun): bash scripts/gpu_check.sh [stage ...]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
run() { # name, timeout, command...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout -k 10 "$t" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
  return $rc
}
: > gpurun_out/summary.txt
stages=${@:-"unfused plain tiles fused rest full smoke bench"}
for s in $stages; do
case $s in
  unfused) run t_unfused 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "unfused or bad_arguments or enumerate" ;;
  plain)   run t_plain 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tiles_plain_io" ;;
  tiles)   run t_tiles 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tiles_only or small_tiles" ;;
  fused)   run t_fused 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused and not unfused" ;;
  rest)    run t_rest 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu ;;
  full)    run t_full 1200 python -m pytest tests/test_gpu_fullsize.py -q -m gpu ;;
  smoke)   run t_smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
  bench)   run bench 900 python bench.py ;;
  benchref) run benchref 600 python bench.py --impl reference --steps 2 --warmup 1 ;;
esac
done
cat gpurun_out/summary.txt
