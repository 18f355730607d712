#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -4
timeout 300 python scripts/time_misc.py > gpurun_out/time_misc.log 2>&1; grep -E "intra32 n" gpurun_out/time_misc.log
timeout 300 python scripts/time_satd.py > gpurun_out/time_satd.log 2>&1; grep -E "imma v2 " gpurun_out/time_satd.log
