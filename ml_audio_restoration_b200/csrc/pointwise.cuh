// Launchers of the CUDA-core kernels (pointwise.cu, lstm.cu).
#pragma once
#include "ar_common.cuh"

namespace ar {

struct DenTailW {
  const float* w0;  // [3 taps][8 chunks][16 cout][4 cin]
  const float* b0;  // [16]
  const float* w1;  // [3][4][8][4]
  const float* b1;  // [8]
  const float* w2;  // [3][8]
  const float* wf;  // [32]
  float b2, bf;
};

int launch_stem(const float* x, int B, int T, int taps, const float* w, const float* bias, const Act& out, int lrelu,
                cudaStream_t stream);
int launch_final_k7(const Act& in, const int* in_coff8, const float* const* w, const float* bias, int nout, float* y,
                    int B, int T, const float* x_lr, cudaStream_t stream);
int launch_den_tail(const Act& f, const float* x, float* y, int B, int T, const DenTailW& w, cudaStream_t stream);
int launch_normalize(float* x, long long n, float target_db, float* scratch, cudaStream_t stream);
int launch_split(const float* audio, long long n, float* chunks, int first, int count, int chunk_size, int overlap,
                 cudaStream_t stream);
int launch_ola(const float* y, float* out, long long n, int n_chunks, int channels, int chunk_size, int overlap, int rate,
               cudaStream_t stream);
int launch_plain_to_c4(const float* x, int B, int C, int T, const Act& out, cudaStream_t stream);
int launch_c4_to_plain(const Act& in, int B, int C, int T, float* y, cudaStream_t stream);

// xp: H8 fp16 [B][256 ch] gate pre-activations (W_ih x + b_ih + b_hh, rows i|f|g|o), whh: [256][64] device,
// h_out: H8 fp16 [B][64 ch]; state_in/out: [B][2][64] (h, c) or nullptr.
int launch_lstm(const Act& xp, const float* whh, const Act& h_out, int B, int T, const float* state_in, float* state_out,
                cudaStream_t stream);

}  // namespace ar
