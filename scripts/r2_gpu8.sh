#!/bin/bash
# round 2, 8-GPU call: host-link ceiling (one process and one process per GPU) and the e2e bench at N = 2, 4, 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(nproc; lscpu | grep -i "model name\|socket\|numa\|^CPU(s)"; free -g | head -2; nvidia-smi topo -m; lspci -tv 2>/dev/null | head -80) > gpurun_out/r2_host_info_n8.log 2>&1
timeout 300 python scripts/time_link_ceiling.py --mb 256 --reps 6 > gpurun_out/r2_link_ceiling_1proc.log 2>&1
for N in 2 4 8; do
  timeout 200 $TR --nproc-per-node $N --master-port $((29500+N)) scripts/time_link_ceiling.py --per-rank --mb 256 --reps 6 2>/dev/null | grep '^{' > gpurun_out/r2_link_ceiling_ranks_n$N.log
done
for N in 8 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --no-secondary > gpurun_out/r2_bench_n$N.log 2> gpurun_out/r2_bench_n$N.err
done
cat gpurun_out/r2_link_ceiling_1proc.log gpurun_out/r2_link_ceiling_ranks_n*.log
for N in 2 4 8; do python - $N <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(f'gpurun_out/r2_bench_n{sys.argv[1]}.log') if l.startswith('{')][-1])
    print(sys.argv[1], d['value'], json.dumps(d['e2e']))
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
done
tail -5 gpurun_out/r2_bench_n8.err
