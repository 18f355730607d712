#!/bin/bash
# N=2 check of the three headline workloads and the reference arm under torchrun
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
timeout 900 $R bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench_full_n2.log 2> gpurun_out/r2b_bench_full_n2.err
timeout 600 $R bench.py --gpus 2 --steps 10 --warmup 3 --workload config3 > gpurun_out/r2b_bench_config3_n2.log 2> gpurun_out/r2b_bench_config3_n2.err
timeout 600 $R bench.py --gpus 2 --steps 10 --warmup 3 --workload config4 > gpurun_out/r2b_bench_config4_n2.log 2> gpurun_out/r2b_bench_config4_n2.err
timeout 600 $R bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/r2b_bench_ref_n2.log 2> gpurun_out/r2b_bench_ref_n2.err
python - <<'PY'
import json
for f in ("full", "config3", "config4", "ref"):
    p = f"gpurun_out/r2b_bench_{f}_n2.log"
    try:
        d = json.loads([l for l in open(p) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'FAILED', e); print(open(p.replace('.log', '.err')).read()[-1500:]); continue
    print(f, d.get('metric'), {k: d.get(k) for k in ('value', 'n_gpus', 'ms_per_step')}, (d.get('e2e') or {}).get('value'), (d.get('e2e_u16_surface') or {}).get('value'))
    for s in d.get('secondary', []): print("   ", s["metric"], f'{s["value"]:.4g}', s["roofline"]["frac"])
PY
