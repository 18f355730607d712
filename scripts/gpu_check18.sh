#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x -k "idct or transpose" 2>&1 | tail -4
timeout 300 python scripts/time_misc.py 2>&1 | grep -E "idct|dct32"
