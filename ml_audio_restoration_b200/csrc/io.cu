// Front end of the restoration path on the GPU: PCM decode + mono mix + sample-rate conversion.
//   load_audio (src/utils/audio_processing.py:10-42): sf.read -> torch.mean(dim=0) -> torchaudio.transforms.Resample
// The resampler restates torchaudio's `sinc_interp_hann` polyphase FIR (lowpass_filter_width 6, rolloff 0.99 -- the
// defaults `Resample(sr, sample_rate)` uses at audio_processing.py:38): with o = orig/gcd, n = new/gcd,
//   base = min(o, n) * rolloff,  width = ceil(6 * o / base),  K = 2*width + o taps,  n phases,
//   h[p][k] = sinc(pi t) * cos^2(pi t / 12) * base / o,   t = clamp((fp32(-p/n) + (k - width)/o) * base, -6, 6)
//   y[m*n + p] = sum_k h[p][k] * x[m*o + k - width]   (x = 0 outside [0, N)),   len(y) = ceil(n * N / o)
// The tap table is built in double precision on the host (as torchaudio does), rounded to fp32 and kept in a small
// per-process cache of device buffers; accumulation is fp32 like the reference's conv1d.
#include <cmath>
#include <map>
#include <mutex>
#include <numeric>
#include <utility>
#include <vector>

#include "ar_common.cuh"

namespace ar {

struct ResampleTable {
  float* dev = nullptr;
  int o = 1, n = 1, width = 0, K = 1;
};

static std::mutex g_tab_mu;
static std::map<std::pair<long long, int>, ResampleTable> g_tabs;   // key: (orig << 32 | new, device)

static int get_table(int orig_sr, int new_sr, const ResampleTable** out) {
  int dev = 0;
  AR_CUDA_OK(cudaGetDevice(&dev));
  const std::pair<long long, int> key(((long long)orig_sr << 32) | (unsigned)new_sr, dev);
  std::lock_guard<std::mutex> lock(g_tab_mu);
  auto it = g_tabs.find(key);
  if (it == g_tabs.end()) {
    ResampleTable t;
    const int g = std::gcd(orig_sr, new_sr);
    t.o = orig_sr / g;
    t.n = new_sr / g;
    const double base = (double)(t.o < t.n ? t.o : t.n) * 0.99;
    t.width = (int)std::ceil(6.0 * t.o / base);
    t.K = 2 * t.width + t.o;
    std::vector<float> h((size_t)t.n * t.K);
    const double pi = 3.14159265358979323846;
    for (int p = 0; p < t.n; ++p)
      for (int k = 0; k < t.K; ++k) {
        // torchaudio forms the phase offset -p/n in float32 before promoting to float64; kept (it moves taps by up to 1e-5)
        const double phase = (double)((float)(-p) / (float)t.n);
        double tt = (phase + (double)(k - t.width) / t.o) * base;
        tt = tt < -6.0 ? -6.0 : (tt > 6.0 ? 6.0 : tt);
        const double c = std::cos(tt * pi / 6.0 / 2.0);
        const double window = c * c;
        const double x = tt * pi;
        const double s = x == 0.0 ? 1.0 : std::sin(x) / x;
        h[(size_t)p * t.K + k] = (float)(s * window * (base / t.o));
      }
    AR_CUDA_OK(cudaMalloc(&t.dev, h.size() * sizeof(float)));
    AR_CUDA_OK(cudaMemcpy(t.dev, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    it = g_tabs.emplace(key, t).first;
  }
  *out = &it->second;
  return AR_OK;
}

long long resample_length(long long n, int orig_sr, int new_sr) {
  const int g = std::gcd(orig_sr, new_sr);
  const long long o = orig_sr / g, nn = new_sr / g;
  return (nn * n + o - 1) / o;   // ceil(new * N / orig)
}

// One output sample per thread; the mono mix (mean over `channels` planar rows) is folded into the tap loop's loads.
// Consecutive threads = consecutive outputs: phase-table rows are read contiguously per thread and the input window
// of neighbouring outputs overlaps almost entirely (L1 hits); K <= a few hundred taps, the table stays in L2.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, int channels, long long n_in,
                                                       const float* __restrict__ h, int o, int n, int width, int K,
                                                       float* __restrict__ y, long long n_out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const long long m = j / n;
  const int p = (int)(j - m * n);
  const float* hp = h + (size_t)p * K;
  const long long i0 = m * o - width;
  const float inv_c = 1.0f / (float)channels;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const long long i = i0 + k;
    if (i < 0 || i >= n_in) continue;
    float v = __ldg(x + i);
    for (int c = 1; c < channels; ++c) v += __ldg(x + (long long)c * n_in + i);
    if (channels > 1) v *= inv_c;
    acc = fmaf(__ldg(hp + k), v, acc);
  }
  y[j] = acc;
}

// plain mono mix (no rate change): y[i] = mean_c x[c][i]
__global__ void mono_mean_kernel(const float* __restrict__ x, int channels, long long n, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  for (int c = 1; c < channels; ++c) v += x[(long long)c * n + i];
  y[i] = channels > 1 ? v / (float)channels : v;   // torch.mean divides the sum
}

// interleaved little-endian PCM16 frames [n][channels] -> planar fp32 [channels][n], x / 32768 (soundfile's float32 read)
__global__ void pcm16_kernel(const short* __restrict__ pcm, int channels, long long n, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < channels; ++c) y[(long long)c * n + i] = (float)pcm[i * channels + c] * (1.0f / 32768.0f);
}

// any WAV sample encoding -> planar fp32 [channels][n], scaled as soundfile's float32 read scales it (integers by
// 2^(bits-1), 8-bit is unsigned with a bias of 128, floats as they are): FMT = AR_PCM_* of include/audiorestore.h
template <int FMT>
__global__ void pcm_decode_kernel(const unsigned char* __restrict__ raw, int channels, long long n, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int c = 0; c < channels; ++c) {
    const long long s = i * channels + c;
    float v;
    if (FMT == AR_PCM_U8) {
      v = ((float)raw[s] - 128.0f) * (1.0f / 128.0f);
    } else if (FMT == AR_PCM_S16) {
      v = (float)reinterpret_cast<const short*>(raw)[s] * (1.0f / 32768.0f);
    } else if (FMT == AR_PCM_S24) {
      const unsigned char* p = raw + 3 * s;                      // packed little-endian, sign in the top byte
      const int w = (int)p[0] | ((int)p[1] << 8) | ((int)(signed char)p[2] << 16);
      v = (float)w * (1.0f / 8388608.0f);
    } else if (FMT == AR_PCM_S32) {
      v = __int2float_rn(reinterpret_cast<const int*>(raw)[s]) * (1.0f / 2147483648.0f);
    } else if (FMT == AR_PCM_F32) {
      v = reinterpret_cast<const float*>(raw)[s];
    } else {
      v = __double2float_rn(reinterpret_cast<const double*>(raw)[s]);
    }
    y[(long long)c * n + i] = v;
  }
}

int launch_resample_mono(const float* x, int channels, long long n, int orig_sr, int new_sr, float* y, long long n_out,
                         cudaStream_t stream) {
  AR_CHECK(x && y && channels >= 1 && n >= 1 && orig_sr >= 1 && new_sr >= 1, AR_ERR_INVALID, "resample: bad argument");
  if (orig_sr == new_sr) {
    AR_CHECK(n_out == n, AR_ERR_INVALID, "resample: output length must equal the input length when the rates match");
    mono_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, channels, n, y);
    AR_CUDA_OK(cudaGetLastError());
    return AR_OK;
  }
  AR_CHECK(n_out == resample_length(n, orig_sr, new_sr), AR_ERR_INVALID, "resample: output length must be ceil(new * n / orig)");
  const ResampleTable* t = nullptr;
  AR_TRY(get_table(orig_sr, new_sr, &t));
  resample_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, stream>>>(x, channels, n, t->dev, t->o, t->n, t->width, t->K, y, n_out);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

int launch_pcm16(const short* pcm, int channels, long long n, float* y, cudaStream_t stream) {
  AR_CHECK(pcm && y && channels >= 1 && n >= 1, AR_ERR_INVALID, "pcm16: bad argument");
  pcm16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pcm, channels, n, y);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

int launch_pcm_decode(const void* raw, int format, int channels, long long n, float* y, cudaStream_t stream) {
  AR_CHECK(raw && y && channels >= 1 && n >= 1, AR_ERR_INVALID, "pcm decode: bad argument");
  const unsigned char* r = reinterpret_cast<const unsigned char*>(raw);
  const unsigned grid = (unsigned)((n + 255) / 256);
  switch (format) {
    case AR_PCM_U8:  pcm_decode_kernel<AR_PCM_U8><<<grid, 256, 0, stream>>>(r, channels, n, y); break;
    case AR_PCM_S16: pcm_decode_kernel<AR_PCM_S16><<<grid, 256, 0, stream>>>(r, channels, n, y); break;
    case AR_PCM_S24: pcm_decode_kernel<AR_PCM_S24><<<grid, 256, 0, stream>>>(r, channels, n, y); break;
    case AR_PCM_S32: pcm_decode_kernel<AR_PCM_S32><<<grid, 256, 0, stream>>>(r, channels, n, y); break;
    case AR_PCM_F32: pcm_decode_kernel<AR_PCM_F32><<<grid, 256, 0, stream>>>(r, channels, n, y); break;
    case AR_PCM_F64: pcm_decode_kernel<AR_PCM_F64><<<grid, 256, 0, stream>>>(r, channels, n, y); break;
    default: AR_CHECK(false, AR_ERR_INVALID, "pcm decode: unknown sample format");
  }
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
