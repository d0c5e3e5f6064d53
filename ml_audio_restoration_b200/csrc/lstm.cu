// Persistent LSTM recurrence (hidden 64, 1 layer, unidirectional) -- stereo_separator.py:37-43,107.
//
// The input projection X*W_ih^T + b_ih + b_hh is a tensor-core GEMM done by the conv engine
// (a k=1 conv, 128 -> 256); this kernel runs only the serial part:
//     g_t = xp_t + W_hh h_{t-1};  c_t = s(f) c_{t-1} + s(i) tanh(g);  h_t = s(o) tanh(c_t)
// One CTA of 256 threads owns S independent sequences for their whole length.  Thread
// (warp w, lane l) owns gate row  (l/8)*64 + 8w + (l%8)  -- its 64 W_hh weights stay in
// registers for the entire kernel -- so the four gates of hidden unit 8w+(l%8) sit in ONE
// warp and are combined with three warp shuffles; h_{t-1} is broadcast from shared memory
// (double buffered => a single block barrier per step).  Everything is fp32; only the copy of
// h that feeds the decoder convs is rounded to TF32.
#include "ar_common.cuh"
#include "pointwise.cuh"
#include <cstdlib>

namespace ar {

constexpr int LSTM_H = 64;
constexpr int LSTM_BLK = 8;  // steps per output flush / input prefetch block

// ex2.approx-based gates: |rel err| ~ 2^-21, far below the TF32 noise of the surrounding convs,
// and ~5x fewer issue slots than expf + IEEE division on the per-step critical path.
__device__ __forceinline__ float ex2_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_f(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_f(1.0f + ex2_f(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - 2.0f * rcp_f(ex2_f(2.8853900817779268f * x) + 1.0f); }

// Packed 2-wide fp32 FMA (Blackwell FFMA2): halves the FMA issue slots of the 64-term dot product.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int S>
__global__ void __launch_bounds__(256, (S == 1) ? 2 : 1)
lstm_kernel(const float* __restrict__ xp, long long xp_bs, int xp_Tp, const float* __restrict__ whh,
            float* __restrict__ hout, long long h_bs, int h_Tp, int B, int T,
            const float* __restrict__ state_in, float* __restrict__ state_out) {
  __shared__ __align__(16) float hbuf[2][S][LSTM_H];
  __shared__ __align__(16) float hstage[2][S][LSTM_BLK][LSTM_H];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gate = lane >> 3;               // 0:i 1:f 2:g 3:o
  const int unit = warp * 8 + (lane & 7);
  const int row = gate * LSTM_H + unit;
  const int seq0 = blockIdx.x * S;

  unsigned long long w2[LSTM_H / 2];  // (w[2k], w[2k+1]) pairs
#pragma unroll
  for (int k = 0; k < LSTM_H; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(whh + row * LSTM_H + k);
    w2[k / 2] = pack2(v.x, v.y);
    w2[k / 2 + 1] = pack2(v.z, v.w);
  }

  float c[S];
  const float* xrow[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int b = min(seq0 + s, B - 1);  // surplus slots replay the last sequence, stores are masked
    c[s] = 0.f;
    float h0 = 0.f;
    if (state_in != nullptr) {
      h0 = state_in[(long long)b * 2 * LSTM_H + unit];
      c[s] = state_in[(long long)b * 2 * LSTM_H + LSTM_H + unit];
    }
    if (gate == 0) hbuf[0][s][unit] = h0;
    xrow[s] = xp + act_off(xp_bs, xp_Tp, b, row >> 2, 0) + (row & 3);
  }
  __syncthreads();

  // Pre-activations are prefetched one 8-step block ahead into a register set that is not live
  // while the current block runs (two sets, ping-pong over a 16-step unrolled body), so the loads
  // have a whole block (~8 steps) to land and never sit on the per-step critical path.
  float xa[S][LSTM_BLK], xb[S][LSTM_BLK];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int k = 0; k < LSTM_BLK; ++k) xa[s][k] = __ldg(xrow[s] + 4 * min(k, T - 1));

  int cur = 0;
  float hlast[S];
#pragma unroll
  for (int s = 0; s < S; ++s) hlast[s] = 0.f;

  auto run_block = [&](float (&xc)[S][LSTM_BLK], float (&xnext)[S][LSTM_BLK], int t0) {
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int k = 0; k < LSTM_BLK; ++k) xnext[s][k] = __ldg(xrow[s] + 4 * min(t0 + LSTM_BLK + k, T - 1));
    const int sb = (t0 / LSTM_BLK) & 1;
#pragma unroll
    for (int k = 0; k < LSTM_BLK; ++k) {
      if (t0 + k < T) {  // uniform across the block
        unsigned long long acc[S][4];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          acc[s][0] = pack2(xc[s][k], 0.f);
          acc[s][1] = 0ull; acc[s][2] = 0ull; acc[s][3] = 0ull;
        }
#pragma unroll
        for (int j = 0; j < LSTM_H; j += 8) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(&hbuf[cur][s][j]);
            const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(&hbuf[cur][s][j + 4]);
            acc[s][0] = fma2(w2[j / 2], h0.x, acc[s][0]);
            acc[s][1] = fma2(w2[j / 2 + 1], h0.y, acc[s][1]);
            acc[s][2] = fma2(w2[j / 2 + 2], h1.x, acc[s][2]);
            acc[s][3] = fma2(w2[j / 2 + 3], h1.y, acc[s][3]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float a0, a1, a2, a3, a4, a5, a6, a7;
          unpack2(acc[s][0], a0, a1);
          unpack2(acc[s][1], a2, a3);
          unpack2(acc[s][2], a4, a5);
          unpack2(acc[s][3], a6, a7);
          const float pre = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
          const float a = (gate == 2) ? tanh_f(pre) : sigmoid_f(pre);
          const float af = __shfl_sync(0xffffffffu, a, (lane & 7) + 8);
          const float ag = __shfl_sync(0xffffffffu, a, (lane & 7) + 16);
          const float ao = __shfl_sync(0xffffffffu, a, (lane & 7) + 24);
          if (gate == 0) {
            c[s] = af * c[s] + a * ag;
            const float h = ao * tanh_f(c[s]);
            hlast[s] = h;
            hbuf[cur ^ 1][s][unit] = h;
            hstage[sb][s][k][unit] = h;
          }
        }
        __syncthreads();
        cur ^= 1;
      }
    }
    // flush this block's hidden states: [16 chunks][<=8 steps] float4 per sequence, coalesced along time
    for (int i = tid; i < S * 16 * LSTM_BLK; i += 256) {
      const int s = i / (16 * LSTM_BLK);
      const int ch = (i / LSTM_BLK) % 16;
      const int k = i % LSTM_BLK;
      const int b = seq0 + s;
      if (b < B && t0 + k < T) {
        const float4 v = *reinterpret_cast<const float4*>(&hstage[sb][s][k][4 * ch]);
        *reinterpret_cast<float4*>(hout + act_off(h_bs, h_Tp, b, ch, t0 + k)) =
            make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      }
    }
  };

  for (int t0 = 0; t0 < T; t0 += 2 * LSTM_BLK) {
    run_block(xa, xb, t0);
    if (t0 + LSTM_BLK < T) run_block(xb, xa, t0 + LSTM_BLK);
  }
  if (state_out != nullptr && gate == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int b = seq0 + s;
      if (b < B) {
        state_out[(long long)b * 2 * LSTM_H + unit] = hlast[s];
        state_out[(long long)b * 2 * LSTM_H + LSTM_H + unit] = c[s];
      }
    }
  }
}

int launch_lstm(const Act& xp, const float* whh, const Act& h_out, int B, int T, const float* state_in, float* state_out,
                cudaStream_t stream) {
  AR_CHECK(T >= 1 && B >= 1, AR_ERR_INVALID, "lstm: empty input");
  // S sequences per CTA.  S=1 runs two CTAs per SM (register-limited); S=2/4 trade thread-level for
  // instruction-level parallelism.  AR_LSTM_S overrides the heuristic (tuning knob).
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("AR_LSTM_S");
    forced = e ? atoi(e) : 0;
  }
  int S = forced ? forced : 1;
#define AR_LSTM_LAUNCH(SS)                                                                                     \
  lstm_kernel<SS><<<(B + SS - 1) / SS, 256, 0, stream>>>(xp.base, xp.bs, xp.Tp, whh, h_out.base, h_out.bs, h_out.Tp, B, T, \
                                                         state_in, state_out)
  if (S == 4) AR_LSTM_LAUNCH(4);
  else if (S == 2) AR_LSTM_LAUNCH(2);
  else AR_LSTM_LAUNCH(1);
#undef AR_LSTM_LAUNCH
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
