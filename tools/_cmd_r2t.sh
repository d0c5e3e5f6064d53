# 16-per-SM LSTM + sub-batched chain: new tests first, then the whole GPU suite, then bench at 16 and 8 chunks per SM
timeout 900 python -m pytest tests/test_gpu_bench_geometry.py -m gpu -q -x -k "sixteen or sub_batched or eight_sequences or full_length or two_halves" > gpurun_out/r2t_newtests.log 2>&1; tail -5 gpurun_out/r2t_newtests.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2t_bench16.log 2>&1; tail -c 1500 gpurun_out/r2t_bench16.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline --chunks-per-step 1184 --batch-chunks 1184 > gpurun_out/r2t_bench8.log 2>&1; tail -c 700 gpurun_out/r2t_bench8.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_tests.log 2>&1; tail -3 gpurun_out/r2t_tests.log
