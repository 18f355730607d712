#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "multi_gpu or device_pointer" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_full_n$N.log 2> gpurun_out/bench_full_n$N.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_full_n$N.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','clocks')}); print(d['roofline']); print(d['e2e'])
for s in d['secondary']: print(s['metric'], f"{s['value']:.4g}", s['roofline']['frac'])
PY
tail -3 gpurun_out/bench_full_n$N.err
