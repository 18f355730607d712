#!/bin/bash
# round 2, second 8-GPU call: the full bench at N = 1, 2, 4, 8 on ONE box (scaling of every secondary incl. config 4 as a CUDA graph,
# intra clocks, e2e forms) + config4 as the headline workload at N = 1 and 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py > gpurun_out/r2_full_n1.log 2> gpurun_out/r2_full_n1.err
for N in 2 4 8; do
  timeout 400 $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N > gpurun_out/r2_full_n$N.log 2> gpurun_out/r2_full_n$N.err
done
timeout 200 python bench.py --workload config4 --steps 30 > gpurun_out/r2_c4_n1.log 2>> gpurun_out/r2_full_n1.err
timeout 300 $TR --nproc-per-node 8 --master-port 29790 bench.py --gpus 8 --workload config4 --steps 30 > gpurun_out/r2_c4_n8.log 2>> gpurun_out/r2_full_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29791 bench.py --gpus 8 --workload config3 --steps 10 > gpurun_out/r2_c3_n8.log 2>> gpurun_out/r2_full_n8.err
python - <<'PY'
import json
def load(f):
    try: return json.loads([l for l in open(f'gpurun_out/{f}.log') if l.startswith('{')][-1])
    except Exception as e: print(f, 'FAILED', e); return None
base = load('r2_full_n1')
for n in (1, 2, 4, 8):
    d = load(f'r2_full_n{n}')
    if not d: continue
    e = d['e2e']
    print(f"N={n} value {d['value']:.4g} (x{d['value']/base['value']:.2f}) e2e {e['value']:.4g} frac {e['roofline']['frac']:.2f} pageable {e['pageable']['value']:.4g} registered {e['registered']['value']:.4g} multi {(e.get('single_process_multi_gpu') or {}).get('value')}")
    bs = {s['metric']: s for s in base['secondary']}
    for s in d['secondary']:
        b = bs.get(s['metric'])
        r = s['value'] / b['value'] if b and b['value'] and s['value'] else None
        print(f"    {s['metric']:42s} {s['value']:.4g}  x{r if r is None else round(r, 2)}  frac {s['roofline']['frac'] and round(s['roofline']['frac'], 3)} {s.get('clocks') or ''}")
for f in ('r2_c4_n1', 'r2_c4_n8', 'r2_c3_n8'):
    d = load(f)
    if d: print(f, d['metric'], f"{d['value']:.5g}", 'e2e', f"{d['e2e']['value']:.4g}", d['gpu_launches'], d['clocks'])
PY
tail -3 gpurun_out/r2_full_n8.err
