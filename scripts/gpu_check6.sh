#!/bin/bash
mkdir -p gpurun_out
./tools/microbench > gpurun_out/microbench.log 2>&1
cat gpurun_out/microbench.log
timeout 600 python scripts/tune_dct.py 64 > gpurun_out/tune_dct2.log 2>&1
cat gpurun_out/tune_dct2.log
ncu --set full --clock-control none --import-source on -k regex:"satd8x8_imma" -s 0 -c 3 \
    -o gpurun_out/prof_satd_imma -f python scripts/profile_kernels.py 4 > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log
