#!/bin/bash
# round 2, GPU call 1 (N=1): parity suite, smoke, host-path mode sweep, link ceiling, bench
mkdir -p gpurun_out
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -15 > gpurun_out/r2_pytest.log
echo "pytest $(( $(date +%s)-S ))s" >> gpurun_out/r2_pytest.log; S=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke $(( $(date +%s)-S ))s" >> gpurun_out/r2_smoke.log; S=$(date +%s)
timeout 600 python scripts/time_e2e_modes.py 16 > gpurun_out/r2_e2e_modes.log 2>&1
echo "modes $(( $(date +%s)-S ))s" >> gpurun_out/r2_e2e_modes.log; S=$(date +%s)
timeout 300 python scripts/time_link_ceiling.py > gpurun_out/r2_link_ceiling_n1.log 2>&1
(nproc; lscpu | head -30; free -g; numactl -H 2>/dev/null; nvidia-smi topo -m) > gpurun_out/r2_host_info.log 2>&1
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err
echo "bench $(( $(date +%s)-S ))s" >> gpurun_out/r2_bench_n1.err
tail -n 6 gpurun_out/r2_pytest.log gpurun_out/r2_smoke.log gpurun_out/r2_bench_n1.err
cat gpurun_out/r2_e2e_modes.log
