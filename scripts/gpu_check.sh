#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, short bench for both DCT variants.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --variant bfly --no-secondary > gpurun_out/bench_bfly.log 2>&1
tail -n 50 gpurun_out/pytest.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench_bfly.log
