#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|FAIL|ALL OK|Error|hazard" gpurun_out/sanitize_$tool.log | head -20
done
