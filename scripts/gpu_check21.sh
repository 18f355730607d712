#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x -k "intra" 2>&1 | tail -3
timeout 300 python scripts/time_misc.py > gpurun_out/time_misc.log 2>&1; grep -E "intra32" gpurun_out/time_misc.log
