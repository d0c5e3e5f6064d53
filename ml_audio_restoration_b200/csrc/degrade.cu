// Synthetic 78 rpm degradation generator on the GPU.
//   simulate_vinyl_artifacts (src/utils/audio_processing.py:122-226): surface noise, Poisson pops with a decaying
//   resonance, high-passed crackle, low-passed rumble, Butterworth roll-off -- the last three through
//   scipy.signal.butter + filtfilt (zero-phase forward/backward IIR, float64) in a Python loop over channels.
// The random draws stay on the host (the Python mirror consumes np.random / torch.randn in the reference's order);
// this file holds the deterministic arithmetic:
//   butter()            scipy.signal.butter(order, wn, 'low'|'high'): analog prototype -> lp2lp / lp2hp -> bilinear -> tf
//   vinyl_mix_kernel    y = audio + noise * level, then every pop (float64 exp / sin, rounded to float32) added in
//                       draw order -- gather form: the thread that owns a sample applies all pops covering it, so
//                       overlapping pops need no atomics and the float32 sum order is the reference's
//   filtfilt            odd extension (float32, like scipy on a float32 row), lfilter_zi initial state, forward pass,
//                       backward pass, all in float64 direct form II transposed.  A recurrence over N samples is
//                       serial; it is made parallel by splitting the extended row into 32-sample sub-blocks:
//                         1. iir_block_kernel<.,false>: every thread runs ITS sub-block from a zero state -> f_k
//                         2. iir_scan_kernel: state at the start of every sub-block, z_{k+1} = A^32 z_k + f_k
//                            (A = the 4x4 state matrix), 256 threads per row each folding a segment, the segment
//                            boundaries chained with A^(32 S) by one thread.  The carry runs in DOUBLE-DOUBLE
//                            (~106-bit) arithmetic: for the 100 Hz rumble low-pass A^32 has entries ~1e4 acting on
//                            states that cancel to ~1e-3, so a float64 carry injects ~1e-15 per sub-block which the
//                            recurrence then amplifies to ~1e-8 (measured: 6e-6 relative); carried in double-double
//                            and rounded once per sub-block the result is within 2e-10 of the serial recurrence
//                         3. iir_block_kernel<.,true>: every thread re-runs its sub-block from the true state and emits
//                       (2 x the flops of the serial loop, N/32-way parallel).  Sub-blocks are staged through shared
//                       memory (pitch 33 doubles: conflict-free) so all global traffic is coalesced.
// HBM-bound / latency-bound CUDA-core work: fp64, ~60 B of traffic per sample per filter.
#include <cmath>
#include <complex>
#include <vector>

#include "ar_common.cuh"
#include "prof.cuh"

namespace ar {

// ----------------------------------------------------------------------------- Butterworth design (host, float64)
int butter(int order, double wn, int highpass, double* b, double* a) {
  AR_CHECK(b && a && order >= 1 && order <= 4, AR_ERR_INVALID, "butter: order must be 1..4");
  AR_CHECK(wn > 0.0 && wn < 1.0, AR_ERR_INVALID, "butter: Digital filter critical frequencies must be 0 < Wn < 1");
  typedef std::complex<double> cd;
  const double pi = 3.141592653589793238462643383279502884;
  std::vector<cd> p(order), z;
  for (int i = 0; i < order; ++i) {                       // buttap: p = -exp(j pi m / 2N), m = -N+1, -N+3, ...
    const double m = -order + 1 + 2 * i;
    p[i] = -std::exp(cd(0.0, pi * m / (2.0 * order)));
  }
  double k = 1.0;
  const double fs = 2.0;
  const double warped = 2.0 * fs * std::tan(pi * wn / fs);
  if (!highpass) {                                        // lp2lp_zpk
    for (auto& v : p) v = warped * v;
    k *= std::pow(warped, order);
  } else {                                                // lp2hp_zpk: zeros at the origin, p -> wo / p
    cd prod(1.0, 0.0);
    for (auto& v : p) prod *= -v;
    k *= (cd(1.0, 0.0) / prod).real();
    for (auto& v : p) v = warped / v;
    z.assign(order, cd(0.0, 0.0));
  }
  const double fs2 = 2.0 * fs;                            // bilinear_zpk
  cd num(1.0, 0.0), den(1.0, 0.0);
  for (auto& v : z) num *= (fs2 - v);
  for (auto& v : p) den *= (fs2 - v);
  std::vector<cd> zz, pz;
  for (auto& v : z) zz.push_back((fs2 + v) / (fs2 - v));
  while ((int)zz.size() < order) zz.push_back(cd(-1.0, 0.0));
  for (auto& v : p) pz.push_back((fs2 + v) / (fs2 - v));
  const double kz = k * (num / den).real();
  auto poly = [&](const std::vector<cd>& roots, double scale, double* out) {   // np.poly: repeated convolution with [1, -r]
    std::vector<cd> c(1, cd(1.0, 0.0));
    for (auto& r : roots) {
      std::vector<cd> nx(c.size() + 1, cd(0.0, 0.0));
      for (size_t i = 0; i < c.size(); ++i) {
        nx[i] += c[i];
        nx[i + 1] -= c[i] * r;
      }
      c.swap(nx);
    }
    for (int i = 0; i <= order; ++i) out[i] = scale * c[i].real();
  };
  poly(zz, kz, b);
  poly(pz, 1.0, a);
  return AR_OK;
}

// ----------------------------------------------------------------------------- pops + surface noise
// Must match ar_pop_t in audiorestore.h.
struct Pop {
  long long loc;
  int length;
  int has_resonance;
  double amp_signed;   // amp * polarity
  double amp;
  double tau;          // sample_rate * decay_time * 0.3
  double omega;        // 2 * pi * resonance_freq
};

constexpr int MIX_THREADS = 256;
constexpr int MIX_PER_THREAD = 4;
constexpr int MIX_TILE = MIX_THREADS * MIX_PER_THREAD;

__global__ void __launch_bounds__(MIX_THREADS) vinyl_mix_kernel(const float* __restrict__ audio, const float* __restrict__ noise,
                                                                 float level, const Pop* __restrict__ pops, int n_pops,
                                                                 double sample_rate, float* __restrict__ y, int rows, long long n) {
  __shared__ int s_hit[MIX_THREADS];
  __shared__ int s_warp[MIX_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tile0 = (long long)blockIdx.x * MIX_TILE;
  const long long tile1 = tile0 + MIX_TILE < n ? tile0 + MIX_TILE : n;
  for (int r = 0; r < rows; ++r)
#pragma unroll
    for (int q = 0; q < MIX_PER_THREAD; ++q) {
      const long long i = tile0 + tid + q * MIX_THREADS;
      if (i < n) y[r * n + i] = __fadd_rn(audio[r * n + i], __fmul_rn(noise[r * n + i], level));
    }
  // pops in draw order, MIX_THREADS candidates per pass; the hits of a pass are compacted in order
  for (int base = 0; base < n_pops; base += MIX_THREADS) {
    const int c = base + tid;
    bool hit = false;
    if (c < n_pops) {
      const long long loc = pops[c].loc;
      hit = loc < tile1 && loc + pops[c].length > tile0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < MIX_THREADS / 32; ++w) {
      before += w < warp ? s_warp[w] : 0;
      total += s_warp[w];
    }
    if (hit) s_hit[before + __popc(bal & ((1u << lane) - 1u))] = c;
    __syncthreads();
    for (int h = 0; h < total; ++h) {
      const Pop p = pops[s_hit[h]];
#pragma unroll
      for (int q = 0; q < MIX_PER_THREAD; ++q) {
        const long long i = tile0 + tid + q * MIX_THREADS;
        const long long k = i - p.loc;
        if (i < n && k >= 0 && k < p.length) {
          const double kd = (double)k;
          const double decay = exp(-kd / p.tau);
          double imp = __dmul_rn(p.amp_signed, decay);
          if (p.has_resonance) {
            const double t = kd / sample_rate;
            const double res = __dmul_rn(__dmul_rn(0.3, sin(__dmul_rn(p.omega, t))), decay);
            imp = __dadd_rn(imp, __dmul_rn(__dmul_rn(res, p.amp), 0.2));
          }
          const float f = (float)imp;
          for (int r = 0; r < rows; ++r) y[r * n + i] = __fadd_rn(y[r * n + i], f);
        }
      }
    }
    __syncthreads();
  }
}

__global__ void vinyl_sum_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ x2,
                                 float* __restrict__ y, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float v = x0[i];
  if (x1) v = __fadd_rn(v, x1[i]);
  if (x2) v = __fadd_rn(v, x2[i]);
  y[i] = v;
}

int launch_vinyl_mix(const float* audio, const float* noise, float level, const void* pops, int n_pops, int sample_rate,
                     float* y, int rows, long long n, cudaStream_t stream) {
  AR_CHECK(audio && noise && y && rows >= 1 && n >= 1 && n_pops >= 0 && (n_pops == 0 || pops) && sample_rate >= 1,
           AR_ERR_INVALID, "vinyl_mix: bad argument");
  vinyl_mix_kernel<<<(unsigned)((n + MIX_TILE - 1) / MIX_TILE), MIX_THREADS, 0, stream>>>(
      audio, noise, level, reinterpret_cast<const Pop*>(pops), n_pops, (double)sample_rate, y, rows, n);
  prof_count_launch(1);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

int launch_vinyl_sum(const float* x0, const float* x1, const float* x2, float* y, long long count, cudaStream_t stream) {
  AR_CHECK(x0 && y && count >= 1, AR_ERR_INVALID, "vinyl_sum: bad argument");
  vinyl_sum_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(x0, x1, x2, y, count);
  prof_count_launch(1);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

// ----------------------------------------------------------------------------- filtfilt
constexpr int IIR_L = 32;                     // samples per sub-block (one thread)
constexpr int IIR_THREADS = 128;              // sub-blocks per CTA
constexpr int IIR_TILE = IIR_L * IIR_THREADS;
constexpr int IIR_PITCH = IIR_L + 1;
constexpr int SCAN_THREADS = 256;

struct IirParams {
  double b[5], a[5];   // normalised by a[0], zero-padded to order 4
  double zi[4];        // lfilter_zi
  double AL[32];       // A^IIR_L          (row major, double-double: [16] high parts then [16] low parts)
  double AS[32];       // A^(IIR_L * S)
  long long n, m;      // row length, extended length n + 2 pad
  int pad, nsub, S;    // S = sub-blocks per scan thread
};

struct IirIn {         // input of the forward pass: (scale * x0) + x1 + x2 in float32, rows of n samples
  const float* x0;
  const float* x1;
  const float* x2;
  float scale;
};

__device__ __forceinline__ float iir_x(const IirIn& in, long long off) {
  float v = __fmul_rn(in.x0[off], in.scale);
  if (in.x1) v = __fadd_rn(v, in.x1[off]);
  if (in.x2) v = __fadd_rn(v, in.x2[off]);
  return v;
}

// odd extension by `pad` samples at both ends, formed in float32 (scipy's odd_ext on a float32 row)
__device__ __forceinline__ float iir_ext(const IirIn& in, long long row_off, long long n, int pad, long long i) {
  if (i < pad) return __fsub_rn(__fmul_rn(2.f, iir_x(in, row_off)), iir_x(in, row_off + (pad - i)));
  i -= pad;
  if (i < n) return iir_x(in, row_off + i);
  return __fsub_rn(__fmul_rn(2.f, iir_x(in, row_off + n - 1)), iir_x(in, row_off + (n - 2 - (i - n))));
}

// one direct-form-II-transposed step in scipy's operation order (separately rounded multiplies and adds)
__device__ __forceinline__ double iir_step(const IirParams& P, double x, double (&z)[4]) {
  const double y = __dadd_rn(z[0], __dmul_rn(P.b[0], x));
  z[0] = __dsub_rn(__dadd_rn(z[1], __dmul_rn(x, P.b[1])), __dmul_rn(y, P.a[1]));
  z[1] = __dsub_rn(__dadd_rn(z[2], __dmul_rn(x, P.b[2])), __dmul_rn(y, P.a[2]));
  z[2] = __dsub_rn(__dadd_rn(z[3], __dmul_rn(x, P.b[3])), __dmul_rn(y, P.a[3]));
  z[3] = __dsub_rn(__dmul_rn(x, P.b[4]), __dmul_rn(y, P.a[4]));
  return y;
}

// double-double (unevaluated sum hi + lo, |lo| <= ulp(hi)/2): error-free transformations with explicit roundings
struct dd {
  double hi, lo;
};
__host__ __device__ __forceinline__ double ar_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
__host__ __device__ __forceinline__ dd dd_fast_two_sum(double a, double b) {   // |a| >= |b|
  const double s = a + b;
  return {s, b - (s - a)};
}
__host__ __device__ __forceinline__ dd dd_add(dd x, dd y) {
  const double s = x.hi + y.hi;
  const double bb = s - x.hi;
  double e = (x.hi - (s - bb)) + (y.hi - bb);
  e += x.lo + y.lo;
  return dd_fast_two_sum(s, e);
}
__host__ __device__ __forceinline__ dd dd_mul(dd x, dd y) {
#ifdef __CUDA_ARCH__
  const double p = __dmul_rn(x.hi, y.hi);   // never contracted into the sum that follows
#else
  const double p = x.hi * y.hi;
#endif
  double e = ar_fma(x.hi, y.hi, -p);
  e += x.hi * y.lo + x.lo * y.hi;
  return dd_fast_two_sum(p, e);
}

// out = f + M z   (M: double-double 4x4, [16] high parts then [16] low parts)
__device__ __forceinline__ void mat4_apply(const double* M, const dd (&z)[4], const double (&f)[4], dd (&out)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    dd acc = {f[i], 0.0};
#pragma unroll
    for (int j = 0; j < 4; ++j) acc = dd_add(acc, dd_mul(dd{M[4 * i + j], M[16 + 4 * i + j]}, z[j]));
    out[i] = acc;
  }
}

// REV = backward pass (input: the forward result read back to front).  EMIT = false: zero-state final state of every
// sub-block -> st;  EMIT = true: start state of every sub-block <- st, outputs written (forward: float64 row of m
// samples; backward: un-reversed, padding stripped, rounded to float32).
template <bool REV, bool EMIT>
__global__ void __launch_bounds__(IIR_THREADS) iir_block_kernel(const __grid_constant__ IirParams P, const IirIn in,
                                                                 const double* __restrict__ yfwd_in, double* __restrict__ yfwd_out,
                                                                 float* __restrict__ yout, double* __restrict__ st) {
  __shared__ double sm[IIR_THREADS * IIR_PITCH];
  const int tid = threadIdx.x;
  const long long row = blockIdx.y;
  const long long tile0 = (long long)blockIdx.x * IIR_TILE;
  for (int e = tid; e < IIR_TILE; e += IIR_THREADS) {
    const long long i = tile0 + e;
    double v = 0.0;
    if (i < P.m) v = REV ? yfwd_in[row * P.m + (P.m - 1 - i)] : (double)iir_ext(in, row * P.n, P.n, P.pad, i);
    sm[(e / IIR_L) * IIR_PITCH + (e % IIR_L)] = v;
  }
  __syncthreads();
  const long long sub = (long long)blockIdx.x * IIR_THREADS + tid;
  const long long left = P.m - sub * IIR_L;
  const int cnt = left >= IIR_L ? IIR_L : (left > 0 ? (int)left : 0);
  if (cnt > 0) {
    double* s = st + (row * P.nsub + sub) * 4;
    double z[4] = {0.0, 0.0, 0.0, 0.0};
    if (EMIT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = s[j];
    }
    double* mine = sm + tid * IIR_PITCH;
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const double y = iir_step(P, mine[k], z);
      if (EMIT) mine[k] = y;
    }
    if (!EMIT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] = z[j];
    }
  }
  if (EMIT) {
    __syncthreads();
    for (int e = tid; e < IIR_TILE; e += IIR_THREADS) {
      const long long i = tile0 + e;
      if (i >= P.m) break;
      const double v = sm[(e / IIR_L) * IIR_PITCH + (e % IIR_L)];
      if (!REV) {
        yfwd_out[row * P.m + i] = v;
      } else {
        const long long p = P.m - 1 - i - P.pad;
        if (p >= 0 && p < P.n) yout[row * P.n + p] = (float)v;
      }
    }
  }
}

// st[row][sub][4]: in = zero-state final state of sub-block `sub`, out = true state at its start.
__global__ void __launch_bounds__(SCAN_THREADS) iir_scan_kernel(const __grid_constant__ IirParams P, int rev, const IirIn in,
                                                                 const double* __restrict__ yfwd, double* __restrict__ st) {
  __shared__ dd F[SCAN_THREADS][4];
  __shared__ dd Z[SCAN_THREADS][4];
  const int tid = threadIdx.x;
  const long long row = blockIdx.x;
  double* f = st + row * P.nsub * 4;
  const long long s0 = (long long)tid * P.S;
  const long long s1 = s0 + P.S < P.nsub ? s0 + P.S : P.nsub;
  dd z[4] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}}, t[4];
  double fv[4];
  for (long long s = s0; s < s1; ++s) {
#pragma unroll
    for (int j = 0; j < 4; ++j) fv[j] = f[s * 4 + j];
    mat4_apply(P.AL, z, fv, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = t[j];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) F[tid][j] = z[j];
  __syncthreads();
  if (tid == 0) {
    // initial state zi * (first sample of the pass): ext[0] forward, the last forward output backward
    const double x0 = rev ? yfwd[row * P.m + (P.m - 1)] : (double)iir_ext(in, row * P.n, P.n, P.pad, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = dd{P.zi[j] * x0, 0.0};
    for (int seg = 0; seg < SCAN_THREADS; ++seg) {
      dd fs[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Z[seg][j] = z[j];
        fs[j] = F[seg][j];
      }
      // z = AS z + F[seg]  (F is itself double-double: add its low parts after the product)
      const double fh[4] = {fs[0].hi, fs[1].hi, fs[2].hi, fs[3].hi};
      mat4_apply(P.AS, z, fh, t);
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = dd_add(t[j], dd{fs[j].lo, 0.0});
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) z[j] = Z[tid][j];
  for (long long s = s0; s < s1; ++s) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      fv[j] = f[s * 4 + j];
      f[s * 4 + j] = z[j].hi;          // normalised: hi is the double nearest to hi + lo
    }
    mat4_apply(P.AL, z, fv, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = t[j];
  }
}

static void mat4_mul(const dd* X, const dd* Y, dd* out) {
  dd r[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      dd acc = {0.0, 0.0};
      for (int k = 0; k < 4; ++k) acc = dd_add(acc, dd_mul(X[4 * i + k], Y[4 * k + j]));
      r[4 * i + j] = acc;
    }
  for (int i = 0; i < 16; ++i) out[i] = r[i];
}

// out[0..15] = high parts, out[16..31] = low parts of A^e (double-double)
static void mat4_pow(const double* A, long long e, double* out) {
  dd base[16], acc[16];
  for (int i = 0; i < 16; ++i) {
    base[i] = dd{A[i], 0.0};
    acc[i] = dd{(i % 5 == 0) ? 1.0 : 0.0, 0.0};
  }
  while (e > 0) {
    if (e & 1) mat4_mul(acc, base, acc);
    mat4_mul(base, base, base);
    e >>= 1;
  }
  for (int i = 0; i < 16; ++i) {
    out[i] = acc[i].hi;
    out[16 + i] = acc[i].lo;
  }
}

static inline long long iir_nsub(long long m) { return (m + IIR_L - 1) / IIR_L; }

int filtfilt_workspace_bytes(int rows, long long n, int order, size_t* bytes) {
  AR_CHECK(bytes && rows >= 1 && n >= 1 && order >= 1 && order <= 4, AR_ERR_INVALID, "filtfilt: bad argument");
  const long long m = n + 6LL * (order + 1);
  *bytes = (size_t)rows * (size_t)(m + 4 * iir_nsub(m)) * sizeof(double) + 256;
  return AR_OK;
}

int launch_filtfilt(const float* x, float scale, const float* add1, const float* add2, float* y, int rows, long long n,
                    const double* b, const double* a, int order, void* ws, size_t ws_bytes, cudaStream_t stream) {
  AR_CHECK(x && y && b && a && rows >= 1 && order >= 1 && order <= 4, AR_ERR_INVALID, "filtfilt: bad argument");
  AR_CHECK(a[0] != 0.0, AR_ERR_INVALID, "filtfilt: a[0] must be non-zero");
  const int pad = 3 * (order + 1);
  if (n <= pad) {
    set_error("filtfilt: The length of the input vector x must be greater than padlen, which is " + std::to_string(pad) + ".");
    return AR_ERR_INVALID;
  }
  size_t need = 0;
  AR_TRY(filtfilt_workspace_bytes(rows, n, order, &need));
  AR_CHECK(ws && ws_bytes >= need, AR_ERR_WORKSPACE, "filtfilt: workspace too small");

  IirParams P = {};
  for (int i = 0; i <= order; ++i) {
    P.b[i] = b[i] / a[0];
    P.a[i] = a[i] / a[0];
  }
  {  // lfilter_zi (scipy 1.18): y_inf = sum(b) / sum(a); zi[k] = zi[k+1] + b[k+1] - y_inf a[k+1]
    double sb = 0.0, sa = 0.0;
    for (int i = 0; i <= order; ++i) {
      sb += P.b[i];
      sa += P.a[i];
    }
    AR_CHECK(sa != 0.0, AR_ERR_INVALID, "filtfilt: filter not stable (sum(a) == 0)");
    const double y_inf = sb / sa;
    double run = 0.0;
    for (int k = order; k >= 1; --k) {
      run += P.b[k] - y_inf * P.a[k];
      P.zi[k - 1] = run;
    }
  }
  double A[16] = {0};
  for (int i = 0; i < 4; ++i) {
    A[4 * i] = -P.a[i + 1];
    if (i < 3) A[4 * i + i + 1] = 1.0;
  }
  P.n = n;
  P.pad = pad;
  P.m = n + 2LL * pad;
  const long long nsub = iir_nsub(P.m);
  AR_CHECK(nsub < (1LL << 31), AR_ERR_INVALID, "filtfilt: row too long");
  P.nsub = (int)nsub;
  P.S = (int)((nsub + SCAN_THREADS - 1) / SCAN_THREADS);
  mat4_pow(A, IIR_L, P.AL);
  mat4_pow(A, (long long)IIR_L * P.S, P.AS);

  const uintptr_t base = (reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255;
  double* const yfwd_all = reinterpret_cast<double*>(base);
  double* const st_all = yfwd_all + (size_t)rows * P.m;
  // rows ride on gridDim.y (<= 65535): larger batches go in slabs
  for (int r0 = 0; r0 < rows; r0 += 65535) {
    const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
    double* yfwd = yfwd_all + (size_t)r0 * P.m;
    double* st = st_all + (size_t)r0 * P.nsub * 4;
    const IirIn in = {x + (size_t)r0 * n, add1 ? add1 + (size_t)r0 * n : nullptr, add2 ? add2 + (size_t)r0 * n : nullptr, scale};
    float* yr = y + (size_t)r0 * n;
    const dim3 grid((unsigned)((P.m + IIR_TILE - 1) / IIR_TILE), (unsigned)nr);
    iir_block_kernel<false, false><<<grid, IIR_THREADS, 0, stream>>>(P, in, nullptr, nullptr, nullptr, st);
    iir_scan_kernel<<<nr, SCAN_THREADS, 0, stream>>>(P, 0, in, nullptr, st);
    iir_block_kernel<false, true><<<grid, IIR_THREADS, 0, stream>>>(P, in, nullptr, yfwd, nullptr, st);
    iir_block_kernel<true, false><<<grid, IIR_THREADS, 0, stream>>>(P, in, yfwd, nullptr, nullptr, st);
    iir_scan_kernel<<<nr, SCAN_THREADS, 0, stream>>>(P, 1, in, yfwd, st);
    iir_block_kernel<true, true><<<grid, IIR_THREADS, 0, stream>>>(P, in, yfwd, nullptr, yr, st);
    prof_count_launch(6);
  }
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
