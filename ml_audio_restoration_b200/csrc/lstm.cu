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

namespace ar {

constexpr int LSTM_H = 64;
constexpr int LSTM_BLK = 8;  // steps per output flush / input prefetch block

// ex2.approx-based gates: |rel err| ~ 2^-21, far below the TF32 noise of the surrounding convs,
// and ~5x fewer issue slots than expf + IEEE division on the per-step critical path.
__device__ __forceinline__ float sigmoid_f(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - 2.0f * __frcp_rn(__expf(2.0f * x) + 1.0f); }

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
__global__ void __launch_bounds__(256, (S <= 2) ? 2 : 1)
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

  float xn[S][LSTM_BLK];  // ring of prefetched pre-activations, each load in flight for 8 steps
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int k = 0; k < LSTM_BLK; ++k) xn[s][k] = __ldg(xrow[s] + 4 * min(k, T - 1));

  int cur = 0;
  float hlast[S];
#pragma unroll
  for (int s = 0; s < S; ++s) hlast[s] = 0.f;
  for (int t0 = 0; t0 < T; t0 += LSTM_BLK) {
    const int sb = (t0 / LSTM_BLK) & 1;
#pragma unroll
    for (int k = 0; k < LSTM_BLK; ++k) {
      if (t0 + k < T) {  // uniform across the block
        unsigned long long acc[S][2];
#pragma unroll
        for (int s = 0; s < S; ++s) {
          acc[s][0] = pack2(xn[s][k], 0.f);
          acc[s][1] = 0ull;
          xn[s][k] = __ldg(xrow[s] + 4 * min(t0 + LSTM_BLK + k, T - 1));  // same slot, 8 steps ahead
        }
#pragma unroll
        for (int j = 0; j < LSTM_H; j += 4) {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const ulonglong2 hv = *reinterpret_cast<const ulonglong2*>(&hbuf[cur][s][j]);
            acc[s][0] = fma2(w2[j / 2], hv.x, acc[s][0]);
            acc[s][1] = fma2(w2[j / 2 + 1], hv.y, acc[s][1]);
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          float a0, a1, a2, a3;
          unpack2(acc[s][0], a0, a1);
          unpack2(acc[s][1], a2, a3);
          const float pre = (a0 + a1) + (a2 + a3);
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
  const int sms = sm_count();
  if (B <= 2 * sms) {
    lstm_kernel<1><<<B, 256, 0, stream>>>(xp.base, xp.bs, xp.Tp, whh, h_out.base, h_out.bs, h_out.Tp, B, T, state_in, state_out);
  } else if (B <= 4 * sms) {
    lstm_kernel<2><<<(B + 1) / 2, 256, 0, stream>>>(xp.base, xp.bs, xp.Tp, whh, h_out.base, h_out.bs, h_out.Tp, B, T, state_in, state_out);
  } else {
    lstm_kernel<4><<<(B + 3) / 4, 256, 0, stream>>>(xp.base, xp.bs, xp.Tp, whh, h_out.base, h_out.bs, h_out.Tp, B, T, state_in, state_out);
  }
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
