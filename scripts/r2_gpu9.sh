#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "search or sad or config4 or smoke" 2>&1 | tail -4 > gpurun_out/r2_pytest9.log
timeout 300 python scripts/time_search.py > gpurun_out/r2_time_search.log 2>&1
cat gpurun_out/r2_pytest9.log gpurun_out/r2_time_search.log
