#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "intra" 2>&1 | tail -8 > gpurun_out/r2_pytest4.log
timeout 300 python scripts/time_intra_modes.py > gpurun_out/r2_intra_modes2.log 2>&1
tail -4 gpurun_out/r2_pytest4.log; cat gpurun_out/r2_intra_modes2.log
