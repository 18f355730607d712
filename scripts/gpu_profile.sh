#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list of the bench command (headline + e2e; the secondary list adds ~1700 back-to-back launches of the
# same kernel for the sustained figure and is profiled per kernel below), the bench-size DCT32 launch with DRAM traffic, and one full capture per kernel.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dct32_imma" -s 4 -c 1 \
    -o gpurun_out/prof_dct32_benchsize -f python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/ncu_benchsize.log 2>&1
# (no --import-source here: with every kernel of the library in one report the embedded sources push it past gpurun's 64 MiB return limit;
#  the SASS-level stall samples the summary uses do not need them)
ncu --set full --clock-control none -k regex:"dct|satd8x8|intra32|sad8x8|quant|tiles_to" -s 0 -c 80 \
    -o gpurun_out/prof_all -f python scripts/profile_kernels.py 16 > gpurun_out/ncu_all.log 2>&1
tail -3 gpurun_out/ncu_all.log
# the per-kernel summary is made here (ncu CLI only); the report itself travels back only if it fits gpurun's 64 MiB return limit
python scripts/ncu_summary.py gpurun_out/prof_all.ncu-rep > gpurun_out/ncu_summary_all.md 2> gpurun_out/ncu_summary_all.err
if [ $(stat -c %s gpurun_out/prof_all.ncu-rep) -gt 50000000 ]; then rm -f gpurun_out/prof_all.ncu-rep; echo "prof_all.ncu-rep dropped (summary kept)"; fi
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
