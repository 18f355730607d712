#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/time_e2e.py > gpurun_out/time_e2e.log 2>&1
cat gpurun_out/time_e2e.log
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -12 gpurun_out/topo.txt; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" 
