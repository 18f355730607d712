#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/time_e2e_modes.py 16 > gpurun_out/r2_e2e_modes2.log 2>&1
timeout 600 python bench.py --no-secondary > gpurun_out/r2_bench_n1b.log 2> gpurun_out/r2_bench_n1b.err
cat gpurun_out/r2_e2e_modes2.log; tail -3 gpurun_out/r2_bench_n1b.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n1b.log') if l.startswith('{')][-1])
print(json.dumps(d['e2e'], indent=1))
PY
