#!/bin/bash
mkdir -p gpurun_out
python scripts/time_membw.py 2>&1 | tee gpurun_out/time_membw.log
