#!/bin/bash
# N-GPU check of the three headline workloads under torchrun: bash scripts/r2b_nN.sh N
N=${1:-2}
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
timeout 900 $R bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2b_bench_full_n$N.log 2> gpurun_out/r2b_bench_full_n$N.err
timeout 600 $R bench.py --gpus $N --steps 10 --warmup 3 --workload config3 > gpurun_out/r2b_bench_config3_n$N.log 2> gpurun_out/r2b_bench_config3_n$N.err
timeout 600 $R bench.py --gpus $N --steps 20 --warmup 3 --workload config4 > gpurun_out/r2b_bench_config4_n$N.log 2> gpurun_out/r2b_bench_config4_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in ("full", "config3", "config4"):
    p = f"gpurun_out/r2b_bench_{f}_n{N}.log"
    try:
        d = json.loads([l for l in open(p) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'FAILED', e); print(open(p.replace('.log', '.err')).read()[-1500:]); continue
    print(f, d.get('metric'), {k: d.get(k) for k in ('value', 'n_gpus', 'ms_per_step')}, (d.get('e2e') or {}).get('value'), (d.get('e2e_u16_surface') or {}).get('value'))
    for s in d.get('secondary', []): print("   ", s["metric"], f'{s["value"]:.4g}', s["roofline"]["frac"])
PY
