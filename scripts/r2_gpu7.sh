#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -8 > gpurun_out/r2_pytest7.log
echo "pytest $(( $(date +%s)-S ))s" >> gpurun_out/r2_pytest7.log; S=$(date +%s)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize.py > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "exit $? after $(( $(date +%s)-S ))s" >> gpurun_out/r2_sanitize_memcheck.log; S=$(date +%s)
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize.py > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo "exit $? after $(( $(date +%s)-S ))s" >> gpurun_out/r2_sanitize_racecheck.log
tail -4 gpurun_out/r2_pytest7.log gpurun_out/r2_sanitize_memcheck.log gpurun_out/r2_sanitize_racecheck.log
grep -c "^ok" gpurun_out/r2_sanitize_memcheck.log gpurun_out/r2_sanitize_racecheck.log; grep "FAIL\|ERROR SUMMARY" gpurun_out/r2_sanitize_*.log | head
