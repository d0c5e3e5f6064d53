// CUDA-core conv engine on H8 (fp16) activations with fp32 accumulation.  Debug cross-check for the
// tcgen05 engines (AR_ENGINE_SIMT): same ConvParams, same packed fp16 weights, same fused epilogue.
#include "ar_common.cuh"

namespace ar {

constexpr int SIMT_NB = 16;  // GEMM columns per thread

__global__ void __launch_bounds__(TILE_M) conv_simt_kernel(const ConvParams p) {
  const int tile = blockIdx.x;
  const int b = tile / p.tiles_per_item;
  const int t = (tile % p.tiles_per_item) * TILE_M + threadIdx.x;
  const int n_base = blockIdx.y * SIMT_NB;
  float acc[SIMT_NB];
#pragma unroll
  for (int i = 0; i < SIMT_NB; ++i) acc[i] = 0.f;

  const int kblocks = p.Cin / 16;
  const int Ns = p.N / p.n_slices;
  const int slice = n_base / Ns, n_loc = n_base % Ns;  // SIMT_NB divides Ns (both multiples of 16)
  for (int kb = 0; kb < kblocks; ++kb) {
    for (int tap = 0; tap < p.taps; ++tap) {
      const int ti = t + tap * p.dil - p.pad_left;
      const bool ok = (ti >= 0) && (ti < p.Tin);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.f;
        if (ok) unpack_half8(*reinterpret_cast<const uint4*>(p.in + act_off(p.in_bs, p.in_Tp, b, p.in_coff8 + kb * 2 + h, ti)), a);
        const uint4* w8 = reinterpret_cast<const uint4*>(p.w) +
                          ((size_t)(((slice * kblocks + kb) * p.taps + tap) * 2 + h) * Ns + n_loc);
#pragma unroll
        for (int i = 0; i < SIMT_NB; ++i) {
          float w[8];
          unpack_half8(__ldg(w8 + i), w);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i] = fmaf(a[j], w[j], acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SIMT_NB / 8; ++c) {
    float rv[8], av[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { rv[i] = 0.f; av[i] = acc[8 * c + i]; }
    if (p.res != nullptr && t < p.Tin)
      unpack_half8(*reinterpret_cast<const uint4*>(p.res + act_off(p.res_bs, p.res_Tp, b, p.res_coff8 + ((n_base + 8 * c) >> 3), t)), rv);
    epilogue_chunk8(p, b, t, n_base + 8 * c, av, rv);
  }
}

int launch_conv_simt(const ConvParams& p, cudaStream_t stream) {
  dim3 grid(p.B * p.tiles_per_item, p.N / SIMT_NB);
  conv_simt_kernel<<<grid, TILE_M, 0, stream>>>(p);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
