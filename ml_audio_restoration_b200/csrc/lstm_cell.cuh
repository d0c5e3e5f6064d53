// Device helpers shared by the LSTM recurrence kernels (lstm.cu, lstm_proj.cu): the pre-scaled cell update, the warp-level
// tensor-core MMA and the 16-byte cp.async.
#pragma once
#include "ar_common.cuh"

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
// Gate pre-activations arrive PRE-SCALED (models.cu scales the rows of W_ih, W_hh and the biases at pack time):
//   p_i, p_f, p_o = log2(e) * (gate pre-activation),   p_g = 2 log2(e) * (cell-candidate pre-activation),
// so every gate function is an ex2 of the negated input with no multiply in front of it.
// One LSTM cell update with 5 ex2 + 2 rcp (the special-function unit, 16 lanes/clk/SM, and the issue slots are the
// throughput limits of the tensor-core recurrence kernels): the three gate functions of the cell state share ONE
// reciprocal,
//   s(f) c + s(i) tanh(g) = [ c (1+b)(1+d) + (1-d)(1+a) ] / [ (1+a)(1+b)(1+d) ],  a=e^-f, b=e^-i, d=e^-2g,
// and so do s(o) tanh(c').  Only the lower side needs a clamp (e^-x overflows for very negative x; for large x it
// underflows to 0, which is exact enough): -20 natural units changes a gate by < 3e-9 and keeps the product of three
// exponentials finite.
constexpr float LSTM_L2E = 1.4426950408889634f;
__device__ __forceinline__ void lstm_cell(float pi, float pf, float pg, float po, float& c, float& h) {
  pi = fmaxf(pi, -20.f * LSTM_L2E);
  pf = fmaxf(pf, -20.f * LSTM_L2E);
  pg = fmaxf(pg, -40.f * LSTM_L2E);
  po = fmaxf(po, -20.f * LSTM_L2E);
  const float a = ex2_f(-pf), b = ex2_f(-pi), d = ex2_f(-pg);
  const float bd = (1.f + b) * (1.f + d);
  const float num = fmaf(c, bd, (1.f - d) * (1.f + a));
  c = num * rcp_f((1.f + a) * bd);
  const float cc = fmaxf(c, -20.f);
  const float q = ex2_f(-po), e = ex2_f(-2.f * LSTM_L2E * cc);
  h = (1.f - e) * rcp_f((1.f + q) * (1.f + e));
}
// Two cells per call with the fp32 adds / multiplies as packed 2-wide instructions (FADD2 / FMUL2 / FFMA2: one issue slot for both
// cells; the special-function calls and the clamps stay scalar).  Operation for operation the same IEEE arithmetic as `lstm_cell`,
// so the results are bit-identical; it only shortens the recurrence warps' instruction stream (the scan kernels are bound by issue
// slots and the special-function unit together).
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// acc + (fp16 pre-activation) in ONE instruction: the mixed-precision add of sm_100 (`add.rn.f32.f16`, SASS FHADD) converts the half
// operand exactly and rounds once -- the same value as __half2float followed by an fp32 add, without the separate conversion.
__device__ __forceinline__ float add_f32_f16(float a, unsigned short h) {
  float r;
  asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(h), "f"(a));
  return r;
}
__device__ __forceinline__ void half2_split(uint32_t v, unsigned short& lo, unsigned short& hi) {
  asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(v));
}
// gate pre-activations of one cell: accumulator (i, f, g, o) + staged fp16 (i, f | g, o)
__device__ __forceinline__ void lstm_preact(float ai, float af, float ag, float ao, uint2 x, float& pi, float& pf, float& pg, float& po) {
  unsigned short xi, xf, xg, xo;
  half2_split(x.x, xi, xf);
  half2_split(x.y, xg, xo);
  pi = add_f32_f16(ai, xi);
  pf = add_f32_f16(af, xf);
  pg = add_f32_f16(ag, xg);
  po = add_f32_f16(ao, xo);
}
__device__ __forceinline__ void lstm_cell2(float pi0, float pi1, float pf0, float pf1, float pg0, float pg1, float po0, float po1,
                                           float (&c)[2], float (&h)[2]) {
  const unsigned long long one = f2_pack(1.f, 1.f);
  const unsigned long long a = f2_pack(ex2_f(-fmaxf(pf0, -20.f * LSTM_L2E)), ex2_f(-fmaxf(pf1, -20.f * LSTM_L2E)));
  const unsigned long long b = f2_pack(ex2_f(-fmaxf(pi0, -20.f * LSTM_L2E)), ex2_f(-fmaxf(pi1, -20.f * LSTM_L2E)));
  const unsigned long long d = f2_pack(ex2_f(-fmaxf(pg0, -40.f * LSTM_L2E)), ex2_f(-fmaxf(pg1, -40.f * LSTM_L2E)));
  const unsigned long long a1 = f2_add(one, a);
  const unsigned long long bd = f2_mul(f2_add(one, b), f2_add(one, d));
  const unsigned long long num = f2_fma(f2_pack(c[0], c[1]), bd, f2_mul(f2_sub(one, d), a1));
  float den0, den1;
  f2_unpack(f2_mul(a1, bd), den0, den1);
  f2_unpack(f2_mul(num, f2_pack(rcp_f(den0), rcp_f(den1))), c[0], c[1]);
  const unsigned long long q = f2_pack(ex2_f(-fmaxf(po0, -20.f * LSTM_L2E)), ex2_f(-fmaxf(po1, -20.f * LSTM_L2E)));
  float ea0, ea1;                                     // -2 log2(e) max(c, -20): the argument of the tanh exponential
  f2_unpack(f2_mul(f2_pack(fmaxf(c[0], -20.f), fmaxf(c[1], -20.f)), f2_pack(-2.f * LSTM_L2E, -2.f * LSTM_L2E)), ea0, ea1);
  const unsigned long long e = f2_pack(ex2_f(ea0), ex2_f(ea1));
  float hd0, hd1;
  f2_unpack(f2_mul(f2_add(one, q), f2_add(one, e)), hd0, hd1);
  f2_unpack(f2_mul(f2_sub(one, e), f2_pack(rcp_f(hd0), rcp_f(hd1))), h[0], h[1]);
}
// pre-scaled variants for the CUDA-core kernel (x already multiplied by log2(e), resp. 2 log2(e))
__device__ __forceinline__ float sigmoid_s(float x) { return rcp_f(1.0f + ex2_f(-x)); }
__device__ __forceinline__ float tanh_s(float x) { return 1.0f - 2.0f * rcp_f(ex2_f(x) + 1.0f); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - 2.0f * rcp_f(ex2_f(2.8853900817779268f * x) + 1.0f); }

constexpr int LM_HS = 80;       // padded row stride of the h exchange buffer (halves): conflict-free 8-byte fragment loads
constexpr int LM_XS = 264;      // padded per-sequence stride of a staged pre-activation row (halves)

__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// W_hh (fp32 [256][64], rows pre-scaled) -> this thread's A fragments of the two 16-row tiles (i|f) and (g|o) of hidden
// units 8 w .. 8 w + 7 (`unit` = 8 w + gid), fp16
__device__ __forceinline__ void load_whh_frags(const float* __restrict__ whh, int unit, int tig, uint32_t (&wfrag)[2][4][4]) {
  auto w2 = [&](int row, int k) {
    const __half2 h = __floats2half2_rn(whh[row * LSTM_H + k], whh[row * LSTM_H + k + 1]);
    return *reinterpret_cast<const uint32_t*>(&h);
  };
#pragma unroll
  for (int tl = 0; tl < 2; ++tl) {
    const int row_lo = (2 * tl) * LSTM_H + unit, row_hi = (2 * tl + 1) * LSTM_H + unit;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      wfrag[tl][kt][0] = w2(row_lo, kt * 16 + 2 * tig);
      wfrag[tl][kt][1] = w2(row_hi, kt * 16 + 2 * tig);
      wfrag[tl][kt][2] = w2(row_lo, kt * 16 + 2 * tig + 8);
      wfrag[tl][kt][3] = w2(row_hi, kt * 16 + 2 * tig + 8);
    }
  }
}
// position of hidden unit `unit` inside its sequence row of the h exchange buffer: within each 16-unit k-tile the pairs
// (2j, 2j+1) and (2j+8, 2j+9) that form one thread's B fragment (b0, b1) are made adjacent => one 8-byte load per k-tile
__device__ __forceinline__ int h_exchange_pos(int unit) {
  return (unit & ~15) + ((((unit & 7) >> 1) * 2 + ((unit >> 3) & 1)) * 2) + (unit & 1);
}

}  // namespace ar
