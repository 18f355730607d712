#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 900 -x -k "satd or dctN" 2>&1 | tail -3
timeout 300 python scripts/time_satd.py 2>&1 | tee gpurun_out/time_satd.log
