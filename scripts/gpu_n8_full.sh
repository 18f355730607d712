#!/bin/bash
# the driver's N=8 line with the secondary section (every rank's shard of config 4, both searches, NCCL frame re-assembly)
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 \
    bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_full_n8.log 2> gpurun_out/bench_full_n8.err
echo "wall $(( $(date +%s)-S ))s rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_full_n8.log') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
    for s in d['secondary']: print("   ", s["metric"], f'{s["value"]:.4g}' if s["value"] is not None else None, s["roofline"]["frac"])
except Exception as e:
    print("FAILED", e); print(open('gpurun_out/bench_full_n8.err').read()[-2500:])
PY
