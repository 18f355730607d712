#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -4
timeout 300 python scripts/time_misc.py > gpurun_out/time_misc.log 2>&1; grep -E "dct4" gpurun_out/time_misc.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|FAIL|ALL OK|hazard" gpurun_out/sanitize_$tool.log | head -8
done
