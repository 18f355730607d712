#!/bin/bash
# N-GPU scaling run (under gpurun --gpus N): reference arm on rank 0, then ours.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 3 --no-secondary > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.log | cut -c1-900
tail -3 gpurun_out/bench_n$N.err
