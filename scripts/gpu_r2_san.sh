#!/bin/bash
# round 2: ONE compute-sanitizer tool per call (B200_PROFILING.md): $1 = memcheck | racecheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/sanitize_cases.py > gpurun_out/r2_san_plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/r2_san_plain.log; exit 1; }
tail -n 1 gpurun_out/r2_san_plain.log
timeout 1500 compute-sanitizer --tool $1 --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/r2_san_$1.log 2>&1; echo "sanitizer $1 rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_cases:|========= (Error|Warning|Race)" gpurun_out/r2_san_$1.log | head -20
