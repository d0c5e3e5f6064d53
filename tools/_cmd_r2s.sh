# round 2, session 2: state of the tree -- GPU tests, bench, launch list with DRAM bytes
python -m pytest tests -m gpu -q -x > gpurun_out/r2s_tests.log 2>&1; tail -3 gpurun_out/r2s_tests.log
python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r2s_bench.log 2>&1; tail -c 600 gpurun_out/r2s_bench.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2s_ncu.log 2>&1
python tools/launch_table.py gpurun_out/r2s_launches.csv > gpurun_out/r2s_table.txt 2>&1; tail -25 gpurun_out/r2s_table.txt
