#!/bin/bash
# default single-GPU bench + a compact print of the line
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
echo "bench wall $(( $(date +%s)-S ))s"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','steps','ms_per_step','gpu_launches','clocks')}); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
for s in d['secondary']: print(s['metric'], f"{s['value']:.4g}", s['roofline']['frac'] and round(s['roofline']['frac'],3), s.get('cpu_baseline') and f"cpu {s['cpu_baseline']['value']:.3g} exact={s['cpu_baseline'].get('gpu_output_bit_exact_on_sample')}")
PY
