#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -15 > gpurun_out/pytest.log
timeout 300 python scripts/time_satd.py > gpurun_out/time_satd.log 2>&1
tail -n 30 gpurun_out/pytest.log gpurun_out/time_satd.log
