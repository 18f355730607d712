#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -6 > gpurun_out/pytest.log
timeout 300 python scripts/time_misc.py > gpurun_out/time_misc.log 2>&1
tail -n 6 gpurun_out/pytest.log; grep -E "decide|dct4" gpurun_out/time_misc.log
