#!/bin/bash
# Full check without the ncu passes: GPU test-suite, smoke, reference arm, default bench (what the driver runs at round end).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -6 > gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err
echo "default bench wall: ${SECONDS}s"
tail -n 5 gpurun_out/pytest_final.log gpurun_out/smoke_final.log; cut -c1-400 gpurun_out/bench_ref_final.log; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_final.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','steps','ms_per_step','gpu_launches','clocks')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
for s in d['secondary']: print(s['metric'], f"{s['value']:.4g}", s['roofline']['frac'] and round(s['roofline']['frac'],3))
PY
