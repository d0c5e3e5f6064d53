export AR_X_NB=2 AR_X_G=1
python tools/_probe_stereo.py stereo 148 44100 > gpurun_out/plain_ncu.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_chain_kernel -s 4 -c 4 -o gpurun_out/r2l_chain -f python tools/_probe_stereo.py stereo 148 44100 > gpurun_out/r2l_ncu.log 2>&1
tail -3 gpurun_out/r2l_ncu.log
