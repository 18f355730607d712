#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/time_misc.py > gpurun_out/time_misc.log 2>&1
tail -n 3 gpurun_out/time_misc.log
