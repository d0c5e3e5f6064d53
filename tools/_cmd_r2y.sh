timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_tests.log 2>&1; tail -3 gpurun_out/r2y_tests.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2y_bench16.log 2>&1; tail -c 700 gpurun_out/r2y_bench16.log
python tools/ws_probe.py 2>&1 | grep chain
