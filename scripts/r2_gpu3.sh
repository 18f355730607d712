#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -15 > gpurun_out/r2_pytest3.log
timeout 300 python scripts/time_encode.py > gpurun_out/r2_time_encode.log 2>&1
timeout 300 python scripts/time_encode.py 2073600 3 >> gpurun_out/r2_time_encode.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:intra32_encode_kernel -c 1 -o gpurun_out/r2_prof_encode -f python scripts/time_encode.py 32400 1 > gpurun_out/r2_ncu_encode.log 2>&1
tail -5 gpurun_out/r2_pytest3.log; cat gpurun_out/r2_time_encode.log
