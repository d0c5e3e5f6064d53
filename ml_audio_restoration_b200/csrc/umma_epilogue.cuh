// Fused epilogue shared by the 1-CTA and 2-CTA tcgen05 conv kernels.
// One warp owns 32 GEMM rows (TMEM lanes = time steps) and `wcols` accumulator columns; per 8-column chunk
// the work is: bias add, LeakyReLU as max(v, slope*v), (residual add), round to fp16 and ONE coalesced 16-byte
// store per thread (consecutive lanes = consecutive rows = consecutive 16 bytes of the H8 layout), plus the
// fused max-pool copy / 2x interleave ("pixel shuffle") / right zero-pad column where the layer needs them.
// Output row addresses are hoisted per tile.
#pragma once
#include "ar_common.cuh"
#include "umma_ptx.cuh"

namespace ar {

// walks the tile pairs of one cluster (pair0, pair0 + step, ...) without divisions in the loop
struct PairIter {
  int b, pi;            // batch item, pair index inside the item
  int step_b, step_p, ppi;
  __device__ PairIter(int pair0, int pair_step, int ppi_) : ppi(ppi_) {
    b = pair0 / ppi_;
    pi = pair0 - b * ppi_;
    step_b = pair_step / ppi_;
    step_p = pair_step - step_b * ppi_;
  }
  __device__ __forceinline__ void next() {
    b += step_b;
    pi += step_p;
    if (pi >= ppi) { pi -= ppi; ++b; }
  }
};

struct EpiRow {
  char* o0;            // output row of this thread, first chunk of this warp's column range
  char* o1;            // interleave mode: the right zero-pad row (written when ok1)
  char* prow;          // pooled row
  const __half* rrow;  // residual row
  long long ostride;   // bytes between consecutive output chunks
  long long pstride;
  long long rstride;   // halves
  bool ok0, ok1, pok, in_ok;
};

template <int MODE, bool POOL, bool RES>
__device__ __forceinline__ EpiRow epi_row(const ConvParams& p, int b, int t, int gcol0) {
  EpiRow r;
  r.in_ok = t < p.Tin;
  r.o1 = nullptr; r.prow = nullptr; r.rrow = nullptr; r.ok1 = false; r.pok = false;
  r.pstride = 0; r.rstride = 0;
  const int tt = r.in_ok ? t : 0;   // keep the address arithmetic in range for masked rows
  r.ostride = (long long)p.out_Tp * 16;     // 8 halves per row
  __half* out = p.out;
  int chunk0;
  if (MODE == MODE_SAME) {
    chunk0 = gcol0 >> 3;
    if (p.out_tblock) {   // time-blocked output (LSTM pre-activations): chunks of one time block are 128 bytes apart
      r.o0 = reinterpret_cast<char*>(out + act_off_tb(p.out_bs, p.N >> 3, b, chunk0, tt));
      r.ostride = 128;
    } else {
      r.o0 = reinterpret_cast<char*>(out + act_off(p.out_bs, p.out_Tp, b, p.out_coff8 + chunk0, tt));
    }
    r.ok0 = r.in_ok && t < p.Tout;
  } else {
    // columns [0,N/2) -> row 2t, [N/2,N) -> row 2t+1; a warp's column range never straddles N/2
    const int hN = p.N >> 1;
    const int phase = gcol0 >= hN;
    chunk0 = (gcol0 - phase * hN) >> 3;
    r.o0 = reinterpret_cast<char*>(out + act_off(p.out_bs, p.out_Tp, b, p.out_coff8 + chunk0, 2 * tt + phase));
    r.ok0 = r.in_ok && (2 * t + phase) < p.Tout;
    // right zero-pad column when the skip tensor is one sample longer (denoiser.py:121-122)
    r.ok1 = (phase == 0) && (t == p.Tin - 1) && (2 * p.Tin < p.Tout);
    r.o1 = reinterpret_cast<char*>(out + act_off(p.out_bs, p.out_Tp, b, p.out_coff8 + chunk0, 2 * p.Tin));
  }
  if (POOL) {
    r.prow = reinterpret_cast<char*>(p.pool + act_off(p.pool_bs, p.pool_Tp, b, p.pool_coff8 + chunk0, tt >> 1));
    r.pstride = (long long)p.pool_Tp * 16;
    r.pok = ((t & 1) == 0) && (t + 1 < p.Tin);
  }
  if (RES) {
    r.rrow = p.res + act_off(p.res_bs, p.res_Tp, b, p.res_coff8 + chunk0, tt);
    r.rstride = (long long)p.res_Tp * 8;
  }
  return r;
}

// residual operand of this warp's (<= 16) columns, fetched BEFORE the accumulator is ready
template <bool RES>
__device__ __forceinline__ void epi_prefetch_res(const EpiRow& r, bool active, uint4 (&resv)[2]) {
  if (RES) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
      resv[c] = (r.in_ok && active) ? __ldcs(reinterpret_cast<const uint4*>(r.rrow + c * r.rstride)) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// ---- the per-element math, written for issue slots (the epilogue warps are instruction-latency bound):
//   bias add and the first LeakyReLU product as packed 2-wide fp32 ops (FADD2 / FMUL2), LeakyReLU as
//   ca*v + cb*|v| with ca = (1+slope)/2, cb = (1-slope)/2 (two FMA-pipe ops instead of FMUL + FMNMX on the half-rate ALU
//   pipe; slope = 1 gives exactly v), and ONE saturating convert for the round-to-fp16 + clamp-to-+-65504
//   (F2FP.SATFINITE.F16.F32.PACK_AB).  2.5 instructions per column instead of 5.75.
__device__ __forceinline__ uint32_t cvt_half2_sat(float lo, float hi) {   // round to nearest, saturate to the finite fp16 range
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// (acc + bias) -> LeakyReLU for two neighbouring columns
__device__ __forceinline__ void bias_act2(uint32_t a0, uint32_t a1, float b0, float b1, float ca, float cb, float& y0, float& y1) {
  unsigned long long v = f2_pack(__uint_as_float(a0), __uint_as_float(a1)), t;
  const unsigned long long bb = f2_pack(b0, b1), cc = f2_pack(ca, ca);
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(v) : "l"(v), "l"(bb));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(v), "l"(cc));
  float v0, v1, t0, t1;
  f2_unpack(v, v0, v1);
  f2_unpack(t, t0, t1);
  y0 = fmaf(fabsf(v0), cb, t0);
  y1 = fmaf(fabsf(v1), cb, t1);
}
// eight accumulator columns + bias -> activation -> (residual) -> packed fp16 row piece
template <bool RES>
__device__ __forceinline__ uint4 epi_chunk8(const uint32_t* a, const float* bias8, float ca, float cb, uint4 resv) {
  const float4 b0 = *reinterpret_cast<const float4*>(bias8);
  const float4 b1 = *reinterpret_cast<const float4*>(bias8 + 4);
  float v[8];
  bias_act2(a[0], a[1], b0.x, b0.y, ca, cb, v[0], v[1]);
  bias_act2(a[2], a[3], b0.z, b0.w, ca, cb, v[2], v[3]);
  bias_act2(a[4], a[5], b1.x, b1.y, ca, cb, v[4], v[5]);
  bias_act2(a[6], a[7], b1.z, b1.w, ca, cb, v[6], v[7]);
  if (RES) {
    float rr[8];
    unpack_half8(resv, rr);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += rr[i];
  }
  return make_uint4(cvt_half2_sat(v[0], v[1]), cvt_half2_sat(v[2], v[3]), cvt_half2_sat(v[4], v[5]), cvt_half2_sat(v[6], v[7]));
}

// Walks `wcols` accumulator columns (16, or a multiple of 32) of this warp's TMEM lanes in blocks of 32 and hands each
// block to f(first column of the block, registers, columns in the block).  The TMEM load of block k + 1 is issued BEFORE
// the math of block k (two register sets), so its latency hides behind that math instead of stalling the warp once per
// block -- ncu showed the epilogue warps of the fused chains waiting on exactly these loads (long_scoreboard, 45 % of the
// stall samples).
// W = columns per block: 32 where registers are plentiful (single-layer engine: one CTA of 320 threads per SM), 16 in the
// fused-chain kernels (600+ threads: two 32-column register sets would spill).
template <int W>
__device__ __forceinline__ void tmem_ld_block(uint32_t taddr, uint32_t (&a)[W]) {
  if (W == 32) tmem_ld32_nowait(taddr, reinterpret_cast<uint32_t(&)[32]>(a));
  else tmem_ld16_nowait(taddr, a);
}
template <int W, class F>
__device__ __forceinline__ void tmem_stream(uint32_t taddr, int wcols, F&& f) {
  uint32_t a0[W], a1[W];
  if (wcols < W) {                       // 16 columns through the 32-wide walker
    tmem_ld16_nowait(taddr, a0);
    tmem_wait_ld();
    f(0, a0, 16);
    return;
  }
  tmem_ld_block<W>(taddr, a0);
  for (int cb = 0; cb < wcols; cb += 2 * W) {
    tmem_wait_ld();
    if (cb + W < wcols) tmem_ld_block<W>(taddr + cb + W, a1);
    f(cb, a0, W);
    if (cb + W < wcols) {
      tmem_wait_ld();
      if (cb + 2 * W < wcols) tmem_ld_block<W>(taddr + cb + 2 * W, a0);
      f(cb + W, a1, W);
    }
  }
}

// one block of W accumulator columns (already in registers) -> bias / activation / residual -> stores of this thread's row
template <int MODE, bool POOL, bool RES, int W>
__device__ __forceinline__ void epi_store_block(const EpiRow& r, const float* s_bias_w, int cb0, const uint32_t (&a)[W], int ncol,
                                                float ca, float cb, const uint4 (&resv)[2]) {
#pragma unroll
  for (int c = 0; c < W / 8; ++c) {
    if (8 * c < ncol) {
      const int ch = (cb0 >> 3) + c;               // 8-column chunk index within this warp's range
      const uint4 packed = epi_chunk8<RES>(a + 8 * c, s_bias_w + cb0 + 8 * c, ca, cb, resv[ch & 1]);
      // activations are written once and read by the NEXT launch, gigabytes later: streaming (evict-first) stores keep the
      // L2 for the rows still to be read (operand prefetch window, residual re-read)
      if (r.ok0) __stcs(reinterpret_cast<uint4*>(r.o0 + (long long)ch * r.ostride), packed);
      if (MODE == MODE_INTERLEAVE2) {
        if (r.ok1) *reinterpret_cast<uint4*>(r.o1 + (long long)ch * r.ostride) = make_uint4(0u, 0u, 0u, 0u);
      }
      if (POOL) {  // MaxPool1d(2,2), floor: rows (t, t+1) live in neighbouring lanes; max of rounded == rounded max
        float q[8], m[8];
        unpack_half8(packed, q);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(q[i], __shfl_down_sync(0xffffffffu, q[i], 1));
        if (r.pok) __stcs(reinterpret_cast<uint4*>(r.prow + (long long)ch * r.pstride), pack_half8(m));
      }
    }
  }
}
template <int MODE, bool POOL, bool RES, int W = 32>
__device__ __forceinline__ void epi_store(const EpiRow& r, const float* s_bias_w /* bias of this warp's first column */, uint32_t taddr,
                                          int wcols, float slope, const uint4 (&resv)[2]) {
  const float ca = 0.5f * (1.0f + slope), cb = 0.5f * (1.0f - slope);
  tmem_stream<W>(taddr, wcols, [&](int cb0, const uint32_t (&a)[W], int ncol) {
    epi_store_block<MODE, POOL, RES, W>(r, s_bias_w, cb0, a, ncol, ca, cb, resv);
  });
}

// (variant, taps) -> kernel instantiation table shared by both engines' launchers
enum EpiVariant { EV_PLAIN = 0, EV_POOL = 1, EV_RES = 2, EV_INTERLEAVE = 3 };
__host__ inline int epi_variant(const ConvParams& p) {
  if (p.mode == MODE_INTERLEAVE2) return EV_INTERLEAVE;
  if (p.pool) return EV_POOL;
  if (p.res) return EV_RES;
  return EV_PLAIN;
}

}  // namespace ar
