mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -6 > gpurun_out/pytest_s3.log
echo "pytest $(( $(date +%s)-S ))s" >> gpurun_out/pytest_s3.log; S=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s3.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_s3.log 2> gpurun_out/bench_s3.err
echo "bench $(( $(date +%s)-S ))s" >> gpurun_out/bench_s3.err
tail -n 4 gpurun_out/pytest_s3.log gpurun_out/smoke_s3.log gpurun_out/bench_s3.err
