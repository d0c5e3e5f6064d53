timeout 600 python -m pytest tests/test_gpu_bench_geometry.py -m gpu -q -x -k "eight_sequences or sub_batched or two_halves or window" > gpurun_out/r2w_newtests.log 2>&1; tail -3 gpurun_out/r2w_newtests.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2w_bench16.log 2>&1; tail -c 900 gpurun_out/r2w_bench16.log
python tools/_probe_stereo.py stereo 2368 8192 > gpurun_out/r2w_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lstm_proj -s 1 -c 1 -o gpurun_out/r2w_lstmproj -f python tools/_probe_stereo.py stereo 2368 8192 > gpurun_out/r2w_ncu.log 2>&1
tail -2 gpurun_out/r2w_ncu.log
