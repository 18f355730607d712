#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/time_search.py > gpurun_out/time_search.log 2>&1
cat gpurun_out/time_search.log
ncu --set full --clock-control none --import-source on -k regex:"dct32_imma|satd8x8" -s 0 -c 40 \
    -o gpurun_out/prof_kernels2 -f python scripts/profile_kernels.py 16 > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log
