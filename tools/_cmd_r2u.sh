# ncu --set full of the 8-sequences-per-CTA LSTM (B = 2368 sequences, T = 8192 steps), source page on
python tools/_probe_stereo.py stereo 2368 8192 > gpurun_out/r2u_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lstm_mmaw -s 1 -c 1 -o gpurun_out/r2u_lstm8 -f python tools/_probe_stereo.py stereo 2368 8192 > gpurun_out/r2u_ncu.log 2>&1
tail -3 gpurun_out/r2u_ncu.log
