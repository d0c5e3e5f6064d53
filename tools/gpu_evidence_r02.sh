#!/bin/bash
# The gpurun commands behind profiles/*_r02.* (each block is one `gpurun --timeout N -- '<block>'` call on one B200).
# A number printed by a run under ncu is never a bench value: the bench lines come from the plain runs.
set -x
# 1. tests + bench line + reference arm
python -m pytest tests -m gpu -q -x
python bench.py > gpurun_out/bench.json                          # -> profiles/bench_r02_final.json
python bench.py --impl reference > gpurun_out/bench_ref.json     # -> profiles/bench_r02_reference_arm.json
# 2. launch list with DRAM bytes of the same bench command (one step), per-launch table, conv-engine DRAM traffic
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches.csv python bench.py --no-secondary --no-cpu-baseline --steps 1 --warmup 1     # -> profiles/launches_r02.csv
python tools/launch_table.py gpurun_out/launches.csv > gpurun_out/table.txt                                      # -> profiles/launch_table_r02.txt
python tools/ncu_traffic.py gpurun_out/launches.csv gpurun_out/conv_traffic.json 2368 2368                       # -> profiles/conv_traffic_r02.json
# 3. ncu --set full of the stereo kernels (after the plain command exited 0)
python tools/probe_forward.py stereo 2368 8192 && ncu --set full --clock-control none --import-source on \
    -k regex:"lstm_proj|conv_chain" -s 8 -c 8 -o gpurun_out/stereo -f python tools/probe_forward.py stereo 2368 8192
python tools/ncu_summary.py gpurun_out/stereo.ncu-rep            # (run where ncu is installed) -> profiles/ncu_stereo_chains_lstm_proj_r02_summary.txt
# 4. scan kernel A/B timing (CUDA events of the library's own profile scopes), pipeline trace of the fused chains, workspace sizes
python tools/probe_scan.py
python tools/chain_trace.py 296 88200
python tools/ws_probe.py
# 5. two GPUs (gpurun --gpus 2)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3   # -> profiles/bench_r02_2gpu.json
# 6. SASS opcode histogram of the shipped library (no GPU needed)
bash tools/sass_histogram.sh > profiles/sass_opcodes_r02.txt
