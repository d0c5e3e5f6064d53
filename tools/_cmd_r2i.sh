export AR_X_NB=2 AR_X_G=1
python tools/ws_probe.py 2>&1 | grep chain > gpurun_out/r2i_ws.log
python tools/chain_trace.py 148 88200 > gpurun_out/r2i_trace_stereo.log 2>&1
AR_LIB_PATH=build/lib_c888.so python tools/chain_trace.py 148 88200 > gpurun_out/r2i_trace_stereo_888.log 2>&1
python tools/chain_trace.py 592 44100 sr > gpurun_out/r2i_trace_sr.log 2>&1
python tools/chain_trace.py 592 44100 denoiser > gpurun_out/r2i_trace_den.log 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/r2i_tests.log 2>&1
python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2i_bench_a.log 2>&1
AR_LIB_PATH=build/lib_c888.so python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2i_bench_b.log 2>&1
AR_LIB_PATH=build/lib_lstm3.so python bench.py --no-secondary --no-cpu-baseline --chunks-per-step 1776 --batch-chunks 1776 > gpurun_out/r2i_bench_c.log 2>&1
cat gpurun_out/r2i_ws.log; grep period gpurun_out/r2i_trace_*.log; tail -2 gpurun_out/r2i_tests.log
