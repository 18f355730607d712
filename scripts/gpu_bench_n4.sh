mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_full_n4.log 2> gpurun_out/bench_full_n4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_full_n4.log') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['roofline']['frac'], d['e2e']['value'])
for s in d['secondary']: print("   ", s["metric"], f'{s["value"]:.4g}' if s["value"] is not None else None)
PY
