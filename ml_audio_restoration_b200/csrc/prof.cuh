// Optional per-category device timing (CUDA events on the launching stream) and a launch counter.
// bench.py uses it to report the dominant kernel's average duration and achieved FLOP/s.
#pragma once
#include <cuda_runtime.h>

namespace ar {
enum ProfCat { CAT_CONV = 0, CAT_LSTM = 1, CAT_STEM = 2, CAT_TAIL = 3, CAT_NORM = 4, CAT_CHUNK = 5, CAT_COUNT = 6 };

void prof_count_launch(int n = 1);
struct ProfScope {  // records start/stop events around the launches issued while it is alive
  ProfScope(int cat, cudaStream_t s, double flops, int launches = 1);
  ~ProfScope();
  int slot;
  cudaStream_t stream;
};
int prof_enable(int on);
int prof_read(double* ms, double* flops, long long* launches, int n);
long long prof_launch_count();
}  // namespace ar
