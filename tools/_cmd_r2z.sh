ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --no-secondary --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2z_ncu.log 2>&1
python tools/launch_table.py gpurun_out/r2z_launches.csv > gpurun_out/r2z_table.txt 2>&1; tail -32 gpurun_out/r2z_table.txt
