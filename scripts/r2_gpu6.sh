#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -15 > gpurun_out/r2_pytest6.log
echo "pytest $(( $(date +%s)-S ))s" >> gpurun_out/r2_pytest6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke6.log 2>&1
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_bench6_n1.log 2> gpurun_out/r2_bench6_n1.err
echo "bench $(( $(date +%s)-S ))s" >> gpurun_out/r2_bench6_n1.err
timeout 300 python bench.py --workload config4 --steps 20 > gpurun_out/r2_bench6_c4.log 2>> gpurun_out/r2_bench6_n1.err
timeout 300 python bench.py --workload config3 --steps 10 > gpurun_out/r2_bench6_c3.log 2>> gpurun_out/r2_bench6_n1.err
tail -n 5 gpurun_out/r2_pytest6.log gpurun_out/r2_smoke6.log gpurun_out/r2_bench6_n1.err
python - <<'PY'
import json
for f in ('r2_bench6_n1','r2_bench6_c4','r2_bench6_c3'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.log') if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'FAILED', e); continue
    print(f, d['metric'], f"{d['value']:.4g}", 'frac', d['roofline'].get('frac'), 'e2e', f"{d['e2e']['value']:.4g}", 'launches', d['gpu_launches'], d['config'].get('workload','')[:60])
    for s in d['secondary']: print('   ', s['metric'], f"{s['value']:.4g}", s['roofline']['frac'] and round(s['roofline']['frac'],3), (s.get('config') or '')[-90:])
PY
