// CUDA-core fp32 conv engine on C4 activations.  Debug cross-check for the tcgen05 engine
// (AR_ENGINE_SIMT); same ConvParams, same packed weights, same fused epilogue.
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

  const int kblocks = p.Cin / 8;
  const int Ns = p.N / p.n_slices;
  const int slice = n_base / Ns, n_loc = n_base % Ns;  // SIMT_NB divides Ns (both multiples of 16)
  for (int kb = 0; kb < kblocks; ++kb) {
    for (int tap = 0; tap < p.taps; ++tap) {
      const int ti = t + tap * p.dil - p.pad_left;
      const bool ok = (ti >= 0) && (ti < p.Tin);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) a = *reinterpret_cast<const float4*>(p.in + act_off(p.in_bs, p.in_Tp, b, p.in_coff4 + kb * 2 + h, ti));
        const float4* w4 = reinterpret_cast<const float4*>(p.w) +
                           ((size_t)(((slice * kblocks + kb) * p.taps + tap) * 2 + h) * Ns + n_loc);
#pragma unroll
        for (int i = 0; i < SIMT_NB; ++i) {
          const float4 w = __ldg(w4 + i);
          acc[i] = fmaf(a.x, w.x, acc[i]);
          acc[i] = fmaf(a.y, w.y, acc[i]);
          acc[i] = fmaf(a.z, w.z, acc[i]);
          acc[i] = fmaf(a.w, w.w, acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SIMT_NB / 4; ++c) {
    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.res != nullptr && t < p.Tin)
      rv = *reinterpret_cast<const float4*>(p.res + act_off(p.res_bs, p.res_Tp, b, p.res_coff4 + ((n_base + 4 * c) >> 2), t));
    epilogue_chunk(p, b, t, n_base + 4 * c, make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]), rv);
  }
}

int launch_conv_simt(const ConvParams& p, cudaStream_t stream) {
  dim3 grid(p.B * p.tiles_per_item, p.N / SIMT_NB);
  conv_simt_kernel<<<grid, TILE_M, 0, stream>>>(p);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
