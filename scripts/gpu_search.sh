mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "search" 2>&1 | tail -15 > gpurun_out/pytest_search.log
timeout 300 python scripts/time_search.py > gpurun_out/time_search.log 2>&1
cat gpurun_out/pytest_search.log gpurun_out/time_search.log
