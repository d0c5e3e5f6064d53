# round-2 evidence: full default bench line (secondary + cpu baseline), reference arm, ncu --set full of the stereo kernels
( time python bench.py > gpurun_out/r2A_bench.json 2> gpurun_out/r2A_bench.err ) 2> gpurun_out/r2A_bench.time; tail -c 400 gpurun_out/r2A_bench.json; cat gpurun_out/r2A_bench.time
( time python bench.py --impl reference > gpurun_out/r2A_bench_ref.json 2> gpurun_out/r2A_bench_ref.err ) 2> gpurun_out/r2A_ref.time; tail -c 600 gpurun_out/r2A_bench_ref.json; cat gpurun_out/r2A_ref.time
python tools/_probe_stereo.py stereo 2368 8192 > gpurun_out/r2A_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lstm_proj|conv_chain" -s 8 -c 8 -o gpurun_out/r2A_stereo -f python tools/_probe_stereo.py stereo 2368 8192 > gpurun_out/r2A_ncu.log 2>&1
tail -2 gpurun_out/r2A_ncu.log
