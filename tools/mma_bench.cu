// Microbenchmark: cycles per tcgen05.mma (kind::f16, K = 16, both operands from shared memory) as a function of
// N, CTA-pair mode and shared-memory layout.  Answers "what does one MMA of the conv engines cost at saturation".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_bench tools/mma_bench.cu && /tmp/mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../ml_audio_restoration_b200/csrc/umma_ptx.cuh"

using namespace ar;

__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lbo, uint32_t sbo, int layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}

// mode: 0 = same A/B every MMA; 1 = A advances like a 3-tap conv stage (tap shift 16 B, K block shift); nacc = accumulators cycled
template <int CTA2>
__global__ void __launch_bounds__(128, 1) mma_bench(int N, int nmma, int layout, int nacc, int a_rows, int a_shift_rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CTA2) tmem_alloc2(smem_u32(&tmem_slot), 512);
    else tmem_alloc(smem_u32(&tmem_slot), 512);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1 && rank == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
    const int Nh = CTA2 ? N / 2 : N;
    const uint32_t idesc = make_idesc_f16(CTA2 ? 256 : 128, N);
    uint64_t ad, bd;
    if (layout == 0) {
      ad = desc_of(a0, a_rows * 16, 128, 0);
      bd = desc_of(b0, Nh * 16, 128, 0);
    } else {
      ad = desc_of(a0, 16, 1024, 2);   // SWIZZLE_128B K-major: SBO = 8 rows x 128 B
      bd = desc_of(b0, 16, 1024, 2);
    }
    if (elect_one()) {
      long long t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        const uint32_t d = tmem_base + (uint32_t)((i % nacc) * N);
        // walk A through a few K blocks like the real kernels do (keeps addresses inside the 96 KB region)
        // a_shift_rows != 0: start the A operand a_shift_rows*(i%7) rows (16 B each) into the run, like tap j of a conv
        const uint64_t adi = ad + (uint64_t)((i % 8) * ((layout == 0 ? 2 * a_rows * 16 : 32) >> 4)) + (uint64_t)((i % 7) * a_shift_rows);
        const uint64_t bdi = bd + (uint64_t)((i % 8) * ((layout == 0 ? 2 * Nh * 16 : 32) >> 4));
        if (CTA2) umma2_f16(d, adi, bdi, idesc, i >= nacc ? 1u : 0u);
        else umma_f16(d, adi, bdi, idesc, i >= nacc ? 1u : 0u);
      }
      long long t1 = clock64();
      if (CTA2) umma_commit2(smem_u32(&bar));
      else umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  } else if (warp == 1 && CTA2) {
    mbar_wait(smem_u32(&bar), 0);   // multicast commit arrives here too
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();
  if (warp == 0) {
    if (CTA2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

template <int CTA2>
static void run(int N, int layout, int nacc, int grid, int a_shift_rows = 0) {
  long long* d;
  cudaMalloc(&d, 16);
  const int nmma = 512, smem = 160 * 1024;
  auto k = mma_bench<CTA2>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, N, nmma, layout, nacc, 136, a_shift_rows, d);
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("run: %s\n", cudaGetErrorString(e)); return; }
  }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double ideal = (CTA2 ? 256.0 : 128.0) * N / (256.0 * (CTA2 ? 2 : 1));
  printf("shift=%d cta_group::%d M=%3d N=%3d layout=%s nacc=%d grid=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma (math floor %.0f)\n",
         a_shift_rows, CTA2 ? 2 : 1, CTA2 ? 256 : 128, N, layout ? "swz128" : "none  ", nacc, grid, (double)h[0] / nmma, (double)h[1] / nmma, ideal);
  cudaFree(d);
}

int main() {
  const int Ns[] = {32, 64, 128, 256};
  for (int layout = 0; layout < 2; ++layout)
    for (int N : Ns) {
      run<0>(N, layout, 1, 148);
      run<1>(N, layout, 1, 148);
    }
  for (int sh : {1, 2, 4, 8}) {   // tap-shifted A start (rows of 16 B): is an unaligned operand start more expensive?
    run<1>(64, 0, 1, 148, sh);
    run<1>(128, 0, 1, 148, sh);
    run<1>(256, 0, 1, 148, sh);
  }
  return 0;
}
