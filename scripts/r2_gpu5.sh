#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_intra_ab.log
for v in old vb vc; do
  echo "== $v" >> gpurun_out/r2_intra_ab.log
  X266_B200_LIB=$PWD/tools/libx266_$v.so timeout 300 python scripts/time_intra_modes.py 2>&1 | grep "i % 35\|mode  0\|mode  6\|mode 22\|mode 26\|mode 30" | head -6 >> gpurun_out/r2_intra_ab.log
done
echo "== new" >> gpurun_out/r2_intra_ab.log
timeout 300 python scripts/time_intra_modes.py 2>&1 | grep "i % 35\|mode  0\|mode  6\|mode 22\|mode 26\|mode 30" | head -6 >> gpurun_out/r2_intra_ab.log
cat gpurun_out/r2_intra_ab.log
