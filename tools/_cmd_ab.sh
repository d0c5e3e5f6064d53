for lib in build/lib_lstm3.so ml_audio_restoration_b200/libaudiorestore_sm100.so; do
  echo "=== $lib"
  export AR_LIB_PATH=$lib AR_X_NB=2 AR_X_G=1
  if [ "$lib" != "build/lib_lstm3.so" ]; then unset AR_X_G; fi
  python tools/chain_trace.py 592 44100 sr 2>&1 | grep period | head -1
  python tools/chain_trace.py 592 44100 denoiser 2>&1 | grep period
  python tools/chain_trace.py 148 88200 2>&1 | grep period
  python bench.py --no-secondary --no-cpu-baseline --steps 2 > gpurun_out/r2q_bench_$(basename $lib).log 2>&1
done
