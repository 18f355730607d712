#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list of the bench command + one full capture per kernel.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dct32_imma|dct32_bfly|satd8x8" -s 0 -c 40 \
    -o gpurun_out/prof_kernels -f python scripts/profile_kernels.py 16 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
