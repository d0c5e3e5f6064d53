// Shared definitions for libaudiorestore_sm100: activation layout, conv parameters,
// the fused conv epilogue and error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/audiorestore.h"

namespace ar {

// ----------------------------------------------------------------------------- errors
void set_error(const std::string& msg);
#define AR_CUDA_OK(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ar::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
      return AR_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)
#define AR_CHECK(cond, code, msg)                                                           \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ar::set_error(msg);                                                                   \
      return (code);                                                                        \
    }                                                                                       \
  } while (0)
#define AR_TRY(expr)                                                                        \
  do {                                                                                      \
    int _r = (expr);                                                                        \
    if (_r != AR_OK) return _r;                                                             \
  } while (0)

// ----------------------------------------------------------------------------- layout
// Internal activations are "C4" channel-blocked:  [B][C/4][Tp][4] fp32, where
//   Tp = HALO + round_up(T, TILE_M) + HALO   rows of 16 bytes,
// row (HALO + t) of chunk c holds channels 4c..4c+3 at time t.  One conv tap of an
// implicit-GEMM tile is then a contiguous run of rows, so a tile is fetched with one bulk
// (TMA) copy per channel chunk and every tap is a 16-byte-granular shift of the UMMA
// shared-memory descriptor.  Rows outside [0,T) hold garbage in HBM; consumers zero them
// in shared memory (conv zero padding), producers never write them.
constexpr int TILE_M = 128;
constexpr int HALO = 8;  // >= max one-sided conv reach: dilation 8 * (3-1)/2
constexpr float LRELU_SLOPE = 0.2f;

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int padded_rows(int T) { return HALO + round_up(T, TILE_M) + HALO; }

struct Act {      // a C4 activation tensor (or a channel window of one)
  float* base = nullptr;  // element (b=0, chunk 0, row 0 == t=-HALO)
  int C = 0;              // channels of the whole buffer
  int T = 0;              // valid length
  int Tp = 0;             // padded rows per chunk
  long long bs = 0;       // floats between batch items = (C/4)*Tp*4
  __host__ __device__ size_t floats_per_item() const { return (size_t)(C / 4) * Tp * 4; }
};

__device__ __forceinline__ long long act_off(long long bs, int Tp, int b, int chunk, int t) {
  return (long long)b * bs + ((long long)chunk * Tp + (HALO + t)) * 4;
}

// ----------------------------------------------------------------------------- conv params
enum ConvMode { MODE_SAME = 0, MODE_INTERLEAVE2 = 1 };

struct ConvParams {
  // input (C4)
  const float* in;
  long long in_bs;
  int in_Tp, in_coff4;   // chunk offset of the first input channel
  int Tin;               // valid input length; GEMM rows are input time positions
  int Cin;               // multiple of 8
  int taps, dil, pad_left;  // tap j reads input row t + j*dil - pad_left
  // weights, packed for the UMMA B operand: [n_slices][Cin/8][taps][2][N/n_slices][4] (tf32-rounded fp32)
  const float* w;
  const float* bias;     // [N]
  int N;                 // GEMM N, multiple of 16, <= 256
  int n_slices;          // column slices the weights are packed in (each is one CTA's resident operand)
  int cta2;              // packed for the 2-CTA engine: slices (2i, 2i+1) are the two halves of pair-slice i
  // output (C4)
  int mode;              // MODE_SAME: out[t]; MODE_INTERLEAVE2: cols [0,N/2)->out[2t], [N/2,N)->out[2t+1]
  float* out;
  long long out_bs;
  int out_Tp, out_coff4;
  int Tout;              // valid output length
  float* pool;           // optional max-pool(2,2) copy of the output (MODE_SAME only)
  long long pool_bs;
  int pool_Tp, pool_coff4;
  const float* res;      // optional residual added after the activation (same geometry as out)
  long long res_bs;
  int res_Tp, res_coff4;
  int lrelu;             // LeakyReLU(0.2) after bias
  int round_tf32;        // round stored values to TF32 (they feed a tensor-core conv)
  int B;
  int tiles_per_item;    // ceil(Tin / TILE_M)
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// Fused epilogue for 4 consecutive GEMM columns [n0, n0+4) of GEMM row t (batch item b).
// Must be called by all 32 lanes of a warp whose lanes hold consecutive rows (the pool
// path exchanges neighbours with shuffles); `acc` is the raw accumulator, `resv` the residual
// operand (ignored unless p.res is set; residual layers are MODE_SAME).
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, int b, int t, int n0, float4 acc, float4 resv) {
  const float4 bias = *reinterpret_cast<const float4*>(p.bias + n0);
  float v[4] = {acc.x + bias.x, acc.y + bias.y, acc.z + bias.z, acc.w + bias.w};
  if (p.lrelu) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = v[i] > 0.f ? v[i] : LRELU_SLOPE * v[i];
  }
  int trow, chunk;
  if (p.mode == MODE_SAME) {
    trow = t;
    chunk = n0 >> 2;
  } else {
    const int half = p.N >> 1;
    const int phase = n0 >= half;
    trow = 2 * t + phase;
    chunk = (n0 - phase * half) >> 2;
  }
  const bool row_ok = (t < p.Tin) && (trow < p.Tout);
  if (p.res != nullptr) {  // residual value supplied by the caller (same row / channels as the output)
    v[0] += resv.x; v[1] += resv.y; v[2] += resv.z; v[3] += resv.w;
  }
  if (p.round_tf32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = to_tf32(v[i]);
  }
  if (row_ok)
    *reinterpret_cast<float4*>(p.out + act_off(p.out_bs, p.out_Tp, b, p.out_coff4 + chunk, trow)) =
        make_float4(v[0], v[1], v[2], v[3]);
  if (p.mode == MODE_INTERLEAVE2 && t == p.Tin - 1 && 2 * p.Tin < p.Tout) {
    // right zero-pad column of the up-sampled half when the skip is one sample longer
    // (denoiser.py:121-122); written once per chunk by the phase-0 call.
    if (n0 < (p.N >> 1))
      *reinterpret_cast<float4*>(p.out + act_off(p.out_bs, p.out_Tp, b, p.out_coff4 + chunk, 2 * p.Tin)) =
          make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (p.pool != nullptr) {  // MaxPool1d(2,2), floor (denoiser.py:18,107)
    float m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = fmaxf(v[i], __shfl_down_sync(0xffffffffu, v[i], 1));
    if (((t & 1) == 0) && (t + 1 < p.Tin))
      *reinterpret_cast<float4*>(p.pool + act_off(p.pool_bs, p.pool_Tp, b, p.pool_coff4 + chunk, t >> 1)) =
          make_float4(m[0], m[1], m[2], m[3]);
  }
}

// ----------------------------------------------------------------------------- launchers
int launch_conv_umma(const ConvParams& p, cudaStream_t stream);
int launch_conv_umma2(const ConvParams& p, cudaStream_t stream);   // cta_group::2 engine (needs p.cta2)
int launch_conv_simt(const ConvParams& p, cudaStream_t stream);
int sm_count();

}  // namespace ar
