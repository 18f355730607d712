#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -15 > gpurun_out/pytest.log
timeout 600 python scripts/tune_dct.py 64 > gpurun_out/tune_dct.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1
tail -n 40 gpurun_out/pytest.log gpurun_out/tune_dct.log; python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','clocks')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); [print(s['metric'], s['value'], s['roofline']['frac']) for s in d['secondary']]"
