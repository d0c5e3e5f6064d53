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

template <int S, int PF>
__global__ void __launch_bounds__(256, (S <= 2) ? 2 : 1)
lstm_kernel(const __half* __restrict__ xp, long long xp_bs, int xp_Tp, const float* __restrict__ whh,
            __half* __restrict__ hout, long long h_bs, int h_Tp, int B, int T,
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
  const __half* xrow[S];
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
    xrow[s] = xp + act_off(xp_bs, xp_Tp, b, row >> 3, 0) + (row & 7);
  }
  __syncthreads();

  // Pre-activations are prefetched PF steps ahead into a register set that is not live while the current
  // PF steps run (two sets, ping-pong over a 2*PF-step unrolled body), so the loads have PF steps to land
  // and never sit on the per-step critical path.
  float xa[S][PF], xb[S][PF];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int k = 0; k < PF; ++k) xa[s][k] = __half2float(__ldg(xrow[s] + 8 * min(k, T - 1)));

  int cur = 0;
  float hlast[S];
#pragma unroll
  for (int s = 0; s < S; ++s) hlast[s] = 0.f;

  auto run_block = [&](float (&xc)[S][PF], float (&xnext)[S][PF], int t0) {
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int k = 0; k < PF; ++k) xnext[s][k] = __half2float(__ldg(xrow[s] + 8 * min(t0 + PF + k, T - 1)));
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int t = t0 + k;
      if (t < T) {  // uniform across the block
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
        const int sb = (t >> 3) & 1;
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
            hstage[sb][s][t & 7][unit] = h;
          }
        }
        __syncthreads();
        cur ^= 1;
        if ((t & 7) == 7 || t == T - 1) {
          // flush up to 8 finished steps: [8 chunks][steps] x 16 bytes per sequence, coalesced along time
          const int tb = t & ~7;
          const int nst = t - tb + 1;
          for (int i = tid; i < S * 8 * LSTM_BLK; i += 256) {
            const int s = i / (8 * LSTM_BLK);
            const int ch = (i / LSTM_BLK) % 8;
            const int kk = i % LSTM_BLK;
            const int b = seq0 + s;
            if (b < B && kk < nst) {
              const float4 v0 = *reinterpret_cast<const float4*>(&hstage[sb][s][kk][8 * ch]);
              const float4 v1 = *reinterpret_cast<const float4*>(&hstage[sb][s][kk][8 * ch + 4]);
              const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
              *reinterpret_cast<uint4*>(hout + act_off(h_bs, h_Tp, b, ch, tb + kk)) = pack_half8(v);
            }
          }
        }
      }
    }
  };

  for (int t0 = 0; t0 < T; t0 += 2 * PF) {
    run_block(xa, xb, t0);
    if (t0 + PF < T) run_block(xb, xa, t0 + PF);
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

// ============================================================================ tensor-core recurrence
// Eight sequences per CTA: the recurrent mat-vec of a step becomes a [256 x 64] x [64 x 8] product run on
// warp-level tensor-core MMAs (mma.sync m16n8k8, TF32 operands, fp32 accumulate).  W_hh is rounded to TF32
// like every other weight of the model and h is rounded to TF32 before it is fed back -- it is the same
// rounded value the decoder convs consume -- while the cell state c, the gate pre-activations and all gate
// math stay fp32 (measured on the oracle: output SNR > 100 dB vs the all-fp32 recurrence, tests/ check it).
// Warp w owns hidden units [8w, 8w+8): its two 16-row MMA tiles hold rows (i,f) and (g,o) of those units, so
// in the accumulator layout one thread ends up with all four gates of ONE unit for TWO sequences and the
// cell update needs no cross-thread exchange.  h goes through shared memory ([seq][unit], stride 68 floats:
// conflict-free both for the update's stores and for the B-fragment loads); one block barrier per step.
constexpr int LM_SEQ = 8;       // sequences per CTA
constexpr int LM_HS = 68;       // padded row stride of the h exchange buffer

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int LM_XS = 264;                        // padded per-sequence stride of a staged pre-activation row (halves)
constexpr int LM_XSTEP = LM_SEQ * LM_XS;          // halves per staged step
constexpr int LM_XBUF = LSTM_BLK * LM_XSTEP;      // halves per 8-step buffer
constexpr int LM_SMEM = 2 * LM_XBUF * 2 + (2 * LM_SEQ * LM_HS + 2 * LSTM_BLK * LM_SEQ * LSTM_H) * 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(256, 1)
lstm_mma_kernel(const __half* __restrict__ xp, long long xp_bs, int xp_Tp, const float* __restrict__ whh,
                __half* __restrict__ hout, long long h_bs, int h_Tp, int B, int T,
                const float* __restrict__ state_in, float* __restrict__ state_out) {
  extern __shared__ __align__(16) float lm_smem[];
  __half* const xs = reinterpret_cast<__half*>(lm_smem);       // [2][8 steps][8 seq][264] fp16: staged gate pre-activations
  float* const hbuf = lm_smem + LM_XBUF;                       // [2][8 seq][68]   (2*LM_XBUF halves == LM_XBUF floats)
  float* const hstage = hbuf + 2 * LM_SEQ * LM_HS;             // [2][8 steps][8 seq][64]
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;        // mma fragment coordinates
  const int unit = warp * 8 + gid;                  // hidden unit whose 4 gates this thread finishes
  const int seq0 = blockIdx.x * LM_SEQ;
  const int sa = 2 * tig, sb2 = 2 * tig + 1;        // the two sequences (columns) this thread finishes

  // A fragments: tile 0 = rows (i | f), tile 1 = rows (g | o) of this warp's 8 units; 8 k-tiles each.
  uint32_t wfrag[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int row_lo = (2 * mt) * LSTM_H + warp * 8 + gid;       // gate i (mt=0) / g (mt=1)
    const int row_hi = (2 * mt + 1) * LSTM_H + warp * 8 + gid;   // gate f / o
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      wfrag[mt][kt][0] = __float_as_uint(to_tf32(whh[row_lo * LSTM_H + kt * 8 + tig]));
      wfrag[mt][kt][1] = __float_as_uint(to_tf32(whh[row_hi * LSTM_H + kt * 8 + tig]));
      wfrag[mt][kt][2] = __float_as_uint(to_tf32(whh[row_lo * LSTM_H + kt * 8 + tig + 4]));
      wfrag[mt][kt][3] = __float_as_uint(to_tf32(whh[row_hi * LSTM_H + kt * 8 + tig + 4]));
    }
  }

  // per-thread state: (unit, seq sa) and (unit, seq sb2)
  const int ba = min(seq0 + sa, B - 1), bb = min(seq0 + sb2, B - 1);
  float c0 = 0.f, c1 = 0.f, hl0 = 0.f, hl1 = 0.f;
  if (state_in != nullptr) {
    c0 = state_in[(long long)ba * 2 * LSTM_H + LSTM_H + unit];
    c1 = state_in[(long long)bb * 2 * LSTM_H + LSTM_H + unit];
    hl0 = state_in[(long long)ba * 2 * LSTM_H + unit];
    hl1 = state_in[(long long)bb * 2 * LSTM_H + unit];
  }
  hbuf[sa * LM_HS + unit] = to_tf32(hl0);
  hbuf[sb2 * LM_HS + unit] = to_tf32(hl1);

  // Staging of the gate pre-activations: 8 steps x 8 sequences x 32 chunks of 16 bytes per block, copied with
  // cp.async (8 pieces per thread, consecutive threads = consecutive steps of one (sequence, chunk) run =>
  // 128-byte coalesced reads), one block ahead of its use.
  const uint32_t xs_u32 = (uint32_t)__cvta_generic_to_shared(xs);
  auto stage_block = [&](int blk) {
    const int t0 = blk * LSTM_BLK;
    const uint32_t dst0 = xs_u32 + (uint32_t)((blk & 1) * LM_XBUF * 2);
#pragma unroll 4
    for (int m = 0; m < 8; ++m) {
      const int i = tid + 256 * m;
      const int k = i & 7, run = i >> 3;
      const int sq = run >> 5, ch = run & 31;
      const int b = min(seq0 + sq, B - 1);
      const int t = min(t0 + k, T - 1);
      cp_async16(dst0 + (uint32_t)((k * LM_XSTEP + sq * LM_XS + ch * 8) * 2), xp + act_off(xp_bs, xp_Tp, b, ch, t));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_block(0);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  int cur = 0;
  const int nblk = (T + LSTM_BLK - 1) / LSTM_BLK;
  for (int blk = 0; blk < nblk; ++blk) {
    if (blk + 1 < nblk) stage_block(blk + 1);
    const __half* xb = xs + (blk & 1) * LM_XBUF;
    float* hst = hstage + (blk & 1) * (LSTM_BLK * LM_SEQ * LSTM_H);
    const int t0 = blk * LSTM_BLK;
    const int nst = min(LSTM_BLK, T - t0);
#pragma unroll
    for (int k = 0; k < LSTM_BLK; ++k) {
      if (k < nst) {  // uniform
        float acc[2][2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][h][i] = 0.f;
        const float* hb = hbuf + cur * (LM_SEQ * LM_HS) + gid * LM_HS + tig;
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
          const uint32_t b0 = __float_as_uint(hb[kt * 8]);
          const uint32_t b1 = __float_as_uint(hb[kt * 8 + 4]);
          mma_tf32_16x8x8(acc[0][kt & 1], wfrag[0][kt], b0, b1);
          mma_tf32_16x8x8(acc[1][kt & 1], wfrag[1][kt], b0, b1);
        }
        const __half* xa = xb + k * LM_XSTEP + sa * LM_XS + unit;
        const __half* xbq = xa + LM_XS;
        // accumulator layout: [0]=(row gid, col 2tig) [1]=(gid, 2tig+1) [2]=(gid+8, 2tig) [3]=(gid+8, 2tig+1)
        const float pi0 = acc[0][0][0] + acc[0][1][0] + __half2float(xa[0]), pi1 = acc[0][0][1] + acc[0][1][1] + __half2float(xbq[0]);
        const float pf0 = acc[0][0][2] + acc[0][1][2] + __half2float(xa[64]), pf1 = acc[0][0][3] + acc[0][1][3] + __half2float(xbq[64]);
        const float pg0 = acc[1][0][0] + acc[1][1][0] + __half2float(xa[128]), pg1 = acc[1][0][1] + acc[1][1][1] + __half2float(xbq[128]);
        const float po0 = acc[1][0][2] + acc[1][1][2] + __half2float(xa[192]), po1 = acc[1][0][3] + acc[1][1][3] + __half2float(xbq[192]);
        c0 = sigmoid_f(pf0) * c0 + sigmoid_f(pi0) * tanh_f(pg0);
        c1 = sigmoid_f(pf1) * c1 + sigmoid_f(pi1) * tanh_f(pg1);
        hl0 = sigmoid_f(po0) * tanh_f(c0);
        hl1 = sigmoid_f(po1) * tanh_f(c1);
        // fed-back h == the fp16-rounded value the decoder convs will read (exactly representable in TF32)
        const float hr0 = __half2float(__float2half_rn(hl0)), hr1 = __half2float(__float2half_rn(hl1));
        float* hn = hbuf + (cur ^ 1) * (LM_SEQ * LM_HS);
        hn[sa * LM_HS + unit] = hr0;
        hn[sb2 * LM_HS + unit] = hr1;
        hst[(k * LM_SEQ + sa) * LSTM_H + unit] = hr0;
        hst[(k * LM_SEQ + sb2) * LSTM_H + unit] = hr1;
        if (k == nst - 1) asm volatile("cp.async.wait_group 0;" ::: "memory");   // next block's staging has landed
        __syncthreads();
        cur ^= 1;
      }
    }
    // flush the block's hidden states: per sequence 8 chunks x steps x 16 bytes, coalesced along time
    for (int i = tid; i < LM_SEQ * 8 * LSTM_BLK; i += 256) {
      const int s = i / (8 * LSTM_BLK);
      const int ch = (i / LSTM_BLK) % 8;
      const int kk = i % LSTM_BLK;
      const int b = seq0 + s;
      if (b < B && kk < nst) {
        const float* src = &hst[(kk * LM_SEQ + s) * LSTM_H + 8 * ch];
        const float4 v0 = *reinterpret_cast<const float4*>(src);
        const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        *reinterpret_cast<uint4*>(hout + act_off(h_bs, h_Tp, b, ch, t0 + kk)) = pack_half8(v);
      }
    }
  }
  if (state_out != nullptr) {
    if (seq0 + sa < B) {
      state_out[(long long)(seq0 + sa) * 2 * LSTM_H + unit] = hl0;
      state_out[(long long)(seq0 + sa) * 2 * LSTM_H + LSTM_H + unit] = c0;
    }
    if (seq0 + sb2 < B) {
      state_out[(long long)(seq0 + sb2) * 2 * LSTM_H + unit] = hl1;
      state_out[(long long)(seq0 + sb2) * 2 * LSTM_H + LSTM_H + unit] = c1;
    }
  }
}

int launch_lstm(const Act& xp, const float* whh, const Act& h_out, int B, int T, const float* state_in, float* state_out,
                cudaStream_t stream) {
  AR_CHECK(T >= 1 && B >= 1, AR_ERR_INVALID, "lstm: empty input");
  // AR_LSTM_S=1|2|4 selects the CUDA-core kernel with S sequences per CTA (cross-check / tuning knob).
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("AR_LSTM_S");
    forced = e ? atoi(e) : 0;
  }
  // Heuristic: the CUDA-core kernel (one sequence per CTA, two CTAs per SM) while that covers the batch;
  // beyond two sequences per SM the tensor-core kernel (eight sequences per CTA) wins.  AR_LSTM_S overrides.
  if (forced == 8 || (forced == 0 && B > 2 * sm_count())) {
    static bool attr_set = false;
    if (!attr_set) {
      AR_CUDA_OK(cudaFuncSetAttribute(lstm_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM));
      attr_set = true;
    }
    lstm_mma_kernel<<<(B + LM_SEQ - 1) / LM_SEQ, 256, LM_SMEM, stream>>>(xp.h(), xp.bs, xp.Tp, whh, h_out.h(), h_out.bs, h_out.Tp, B, T,
                                                                  state_in, state_out);
    AR_CUDA_OK(cudaGetLastError());
    return AR_OK;
  }
  int S = forced ? forced : 1;
#define AR_LSTM_LAUNCH(SS, PF)                                                                                 \
  lstm_kernel<SS, PF><<<(B + SS - 1) / SS, 256, 0, stream>>>(xp.h(), xp.bs, xp.Tp, whh, h_out.h(), h_out.bs, h_out.Tp, B, \
                                                             T, state_in, state_out)
  if (S == 4) AR_LSTM_LAUNCH(4, 4);
  else if (S == 2) AR_LSTM_LAUNCH(2, 4);
  else AR_LSTM_LAUNCH(1, 8);
#undef AR_LSTM_LAUNCH
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
