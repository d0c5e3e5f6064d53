// Microbenchmark: cost of an mbarrier wait that is already satisfied, a tcgen05.commit, and an elect+MMA issue burst.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ml_audio_restoration_b200/csrc/umma_ptx.cuh"
using namespace ar;

__global__ void bar_bench(long long* out) {
  __shared__ __align__(8) unsigned long long bars[16];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) for (int i = 0; i < 16; ++i) mbar_arrive(smem_u32(&bars[i]));   // complete phase 0 of all
  __syncthreads();
  if (threadIdx.x < 32) {
    // (a) whole warp waits on completed barriers
    long long t0 = clock64();
    for (int i = 0; i < 256; ++i) mbar_wait(smem_u32(&bars[i & 15]), 0);
    long long t1 = clock64();
    // (b) one lane polls, then __syncwarp
    for (int i = 0; i < 256; ++i) {
      if (lane == 0) mbar_wait(smem_u32(&bars[i & 15]), 0);
      __syncwarp();
    }
    long long t2 = clock64();
    // (c) raw try_wait without the spin bookkeeping
    uint32_t acc = 0;
    for (int i = 0; i < 256; ++i) {
      uint32_t done;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bars[i & 15])), "r"(0) : "memory");
      acc += done;
    }
    long long t3 = clock64();
    // (d) test_wait
    for (int i = 0; i < 256; ++i) {
      uint32_t done;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bars[i & 15])), "r"(0) : "memory");
      acc += done;
    }
    long long t4 = clock64();
    // (e) dependent chain: wait -> tcgen05.fence -> elect
    for (int i = 0; i < 256; ++i) {
      mbar_wait(smem_u32(&bars[i & 15]), 0);
      tc_fence_after();
      if (elect_one()) acc += 1;
      __syncwarp();
    }
    long long t5 = clock64();
    if (lane == 0) {
      out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = acc;
    }
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int r = 0; r < 2; ++r) { bar_bench<<<1, 64>>>(d); cudaDeviceSynchronize(); }
  long long h[6]; cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("per op (cycles): warp mbar_wait %.1f | lane0 wait+syncwarp %.1f | raw try_wait %.1f | test_wait %.1f | wait+fence+elect %.1f\n",
         h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0);
  return 0;
}
