// Throughput of the legacy warp-level tensor-core path (mma.sync m16n8k16 f16 -> f32, SASS HMMA.16816.F32) on sm_100a:
// cycles per instruction per SM sub-partition as a function of warps per SM.  The LSTM recurrence (lstm.cu) issues 16 of
// these per sequence step.      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/hmma_bench tools/hmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void hmma_loop(int iters, float* out, long long* cycles) {
  float d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x3c003c00u, b1 = 0x38003800u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {       // warps per SM (one CTA per SM)
    hmma_loop<<<sms, warps * 32>>>(100, out, cyc);
    cudaDeviceSynchronize();
    hmma_loop<<<sms, warps * 32>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_warp = (double)h / (iters * 8.0);                       // cycles per HMMA seen by one warp
    const double per_smsp = per_warp / (warps < 4 ? 1.0 : warps / 4.0);      // cycles per HMMA per sub-partition
    printf("warps/SM %2d: %.2f cycles per HMMA per warp, %.2f cycles per HMMA per SM sub-partition (%.0f MAC/clk/SM)\n", warps, per_warp,
           per_smsp, 2048.0 * (warps < 4 ? warps : 4) / per_smsp);
  }
  return 0;
}
