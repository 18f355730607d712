#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_full_n2.log 2> gpurun_out/bench_full_n2.err

python - <<'PY'
import json
for f in ("gpurun_out/bench_full_n2.log",):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'FAILED', e); print(open(f.replace('.log','.err')).read()[-1500:]); continue
    print(f, {k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline'])
    for s in d['secondary']: print("   ", s["metric"], f'{s["value"]:.4g}', s["roofline"]["frac"])
PY
