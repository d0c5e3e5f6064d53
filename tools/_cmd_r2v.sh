# fused projection + recurrence kernel: the tests that reach it, under a timeout (a wedged pipeline traps after ~2 s)
timeout 600 python -m pytest tests/test_gpu_bench_geometry.py -m gpu -q -x -k "eight_sequences or sub_batched or two_halves or window" > gpurun_out/r2v_newtests.log 2>&1; tail -15 gpurun_out/r2v_newtests.log
timeout 600 python -m pytest tests/test_gpu_bench_geometry.py -m gpu -q -x -k "sixteen or full_length" > gpurun_out/r2v_newtests2.log 2>&1; tail -5 gpurun_out/r2v_newtests2.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2v_bench16.log 2>&1; tail -c 1800 gpurun_out/r2v_bench16.log
