// The 32 -> 1 k7 output heads (super_resolution.py:62 `reconstruction` + the F.interpolate residual of :96-99;
// stereo_separator.py:81 `{left,right}_decoder.9`) on the tensor core.
//
// A Cout = 1 conv as an implicit GEMM would waste the whole N dimension (one useful column, 7 MMAs per K block).
// Instead the SEVEN TAPS go along N: for the un-shifted activation rows x[t] one MMA per K block computes
//     D[t][j] = sum_c x[t][c] * W_j[c]            (j = 0..6, padded to 8 columns per head)
// and the conv is a shifted sum over accumulator rows,  y[u] = b + sum_j D[u + j - 3][j].  A 128-row tile therefore
// yields 122 outputs (tile stride 122) from K/16 = 2 MMAs (4 for the two stereo heads, whose 32-channel inputs sit side
// by side in one 64-channel tensor and whose taps take columns 0..6 and 8..14) instead of 14 (28).  The shifted sum runs
// through shared memory: every epilogue thread parks its row's 16 accumulator columns, the four epilogue warps meet at a
// named barrier, and thread r adds the 7 (x2) diagonal entries of rows r..r+6.
// HBM-bound by construction: 64 B (128 B) of fp16 activations in, 4 B (8 B) of fp32 out per output sample.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ar_common.cuh"
#include "pointwise.cuh"
#include "umma_ptx.cuh"
#include "umma_epilogue.cuh"

namespace ar {

constexpr int FU_THREADS = 192;              // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int FU_OUT = TILE_M - 6;           // outputs per tile
constexpr int FU_STAGES = 8;
constexpr int FU_RING_BYTES = 4 * TILE_M * 2 * 4;
constexpr int FU_SROW = 17;                  // floats per parked accumulator row (16 + 1: conflict-free diagonal reads)

struct FinalUmmaArgs {
  const __half* in;       // H8 activations, nheads*32 channels from chunk in_coff8
  long long in_bs;
  int in_Tp, in_coff8;
  const __half* w;        // packed B operand: [K/16][2][16][8] fp16
  float bias[2];
  float* y;               // [B][nheads][T] fp32
  const float* x_lr;      // optional [B][T/2]: linear x2 interpolation residual (super-resolution)
  int nheads, T, B, tiles_per_item;
};

__global__ void __launch_bounds__(FU_THREADS, 1) final_umma_kernel(const __grid_constant__ FinalUmmaArgs a, int num_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int K = a.nheads * 32;
  const int n_chunks = K >> 3;
  const int stage_bytes = n_chunks * TILE_M * 16;
  const int w_bytes = (K >> 4) * 2 * 16 * 16;
  // shared memory: [weights][FU_STAGES activation stages][parked rows 2 x 128 x 17 floats][barriers]
  const uint32_t w_base = smem_u32(smem);
  const uint32_t stage_base = w_base + 2048;
  uint8_t* const stage_ptr = smem + 2048;
  float* const park = reinterpret_cast<float*>(smem + 2048 + FU_STAGES * stage_bytes);
  float* const ring = park + 2 * TILE_M * FU_SROW;                      // [4 tiles][128 rows][2] low-rate samples (interp residual)
  const uint32_t bar_base = stage_base + FU_STAGES * stage_bytes + 2 * TILE_M * FU_SROW * 4 + FU_RING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (FU_STAGES + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * FU_STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * FU_STAGES + 2 + i); };
  const uint32_t w_bar = bar_base + 8u * (2 * FU_STAGES + 4);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
      smem + 2048 + FU_STAGES * stage_bytes + 2 * TILE_M * FU_SROW * 4 + FU_RING_BYTES + 8 * (2 * FU_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < FU_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 4);
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tpi = a.tiles_per_item;
  const int tile0 = blockIdx.x, tile_step = gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (elect_one()) {
      mbar_expect_tx(w_bar, (uint32_t)w_bytes);
      bulk_g2s(w_base, a.w, (uint32_t)w_bytes, w_bar);
      int s = 0;
      uint32_t ph = 0;
      const long long chunk_stride = (long long)a.in_Tp * 8;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int b = tile / tpi;
        const int tstart = (tile - b * tpi) * FU_OUT - 3;                     // first row of the tile (may be -3)
        int rows = a.in_Tp - (HALO + tstart);                                  // rows that exist in the buffer from there
        if (rows > TILE_M) rows = TILE_M;
        const __half* src = a.in + act_off(a.in_bs, a.in_Tp, b, a.in_coff8, tstart);
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)(n_chunks * rows * 16));
        uint32_t dst = stage_base + s * stage_bytes;
        for (int c = 0; c < n_chunks; ++c) {
          bulk_g2s(dst, src, (uint32_t)(rows * 16), full_bar(s));
          dst += TILE_M * 16;
          src += chunk_stride;
        }
        if (++s == FU_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA warp
    const uint32_t idesc = make_idesc_f16(128, 16);
    const uint64_t a_desc_hi = make_desc(0u, (uint32_t)(TILE_M * 16), 128u);
    const uint64_t b_desc_hi = make_desc(0u, 16u * 16u, 128u);
    mbar_wait(w_bar, 0);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int b = tile / tpi;
      const int tstart = (tile - b * tpi) * FU_OUT - 3;
      const bool edge = tstart < 0 || tstart + TILE_M > a.T;
      const int buf = it & 1;
      mbar_wait(tempty_bar(buf), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      mbar_wait(full_bar(s), ph);
      if (edge) {   // conv zero padding (and rows of the run that do not exist)
        uint8_t* a_ptr = stage_ptr + s * stage_bytes;
        for (int r = lane; r < TILE_M; r += 32) {
          const int t = tstart + r;
          if (t < 0 || t >= a.T)
            for (int c = 0; c < n_chunks; ++c)
              *reinterpret_cast<float4*>(a_ptr + (c * TILE_M + r) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_async_smem();
        __syncwarp();
      }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 16);
        uint32_t a_addr = (stage_base + s * stage_bytes) >> 4;
        uint32_t b_addr = w_base >> 4;
        for (int kb = 0; kb < (K >> 4); ++kb) {
          umma_f16(d_tmem, a_desc_hi | (uint64_t)a_addr, b_desc_hi | (uint64_t)b_addr, idesc, kb ? 1u : 0u);
          a_addr += 2 * TILE_M;        // two 8-channel chunks of 128 rows x 16 B, in 16-byte units
          b_addr += 2 * 16;            // [2][16][8] halves = 512 B
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(buf));
      }
      __syncwarp();
      if (++s == FU_STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: park rows, barrier, shifted sum
    const int q = warp & 3;                      // TMEM lane quarter
    const int r = q * 32 + lane;                 // accumulator row = output index inside the tile
    // Linear x2 interpolation residual (App. B.3): the two low-rate samples an output needs are fetched with 4-byte
    // cp.async TWO TILES AHEAD into a private slot of a small ring (no cross-thread sharing, so a wait_group suffices).
    // Loaded in the iteration that uses them they put a full global-load latency (~1000 cycles) on every tile.
    const bool interp = a.x_lr != nullptr;
    const uint32_t ring_u32 = smem_u32(ring);
    auto prefetch_lr = [&](int tile, int slot) {
      if (interp && tile < num_tiles) {
        const int b = tile / tpi;
        const int u = (tile - b * tpi) * FU_OUT + r;
        if (r < FU_OUT && u < a.T) {
          const int Tl = a.T >> 1;
          const float* xl = a.x_lr + (long long)b * Tl;
          const int s2 = u >> 1;
          const int nb = (u & 1) ? (s2 + 1 < Tl ? s2 + 1 : Tl - 1) : (s2 > 0 ? s2 - 1 : 0);
          const uint32_t dst = ring_u32 + (uint32_t)((slot * TILE_M + r) * 8);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(xl + s2) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u), "l"(xl + nb) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch_lr(tile0, 0);
    prefetch_lr(tile0 + tile_step, 1);
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int b = tile / tpi;
      const int u = (tile - b * tpi) * FU_OUT + r;          // output time of this thread
      prefetch_lr(tile + 2 * tile_step, (it + 2) & 3);
      const int buf = it & 1;
      float* const pk = park + buf * (TILE_M * FU_SROW);
      mbar_wait(tfull_bar(buf), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 16), acc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));          // accumulator drained: the MMA warp may overwrite it
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[r * FU_SROW + i] = __uint_as_float(acc[i]);
      asm volatile("bar.sync 1, 128;" ::: "memory");        // all 128 rows parked (the four epilogue warps only)
      if (r < FU_OUT && u < a.T) {
        float y0 = a.bias[0], y1 = a.bias[1];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          y0 += pk[(r + j) * FU_SROW + j];
          if (a.nheads == 2) y1 += pk[(r + j) * FU_SROW + 8 + j];
        }
        if (interp) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");           // this tile's pair has landed (two newer groups may fly)
          const float2 lr = *reinterpret_cast<const float2*>(ring + ((it & 3) * TILE_M + r) * 2);
          y0 += 0.75f * lr.x + 0.25f * lr.y;                               // x[s] weight 0.75, the neighbour 0.25 (edge-clamped)
        }
        float* yo = a.y + (long long)b * a.nheads * a.T + u;
        yo[0] = y0;
        if (a.nheads == 2) yo[a.T] = y1;
      }
      // the next tile parks into the other buffer; the barrier of that tile orders this tile's reads before the
      // writes of the tile after it
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 32);
}

// B operand: K-major no-swizzle [K/16][2 (k halves)][16 columns][8 k] fp16; column h*8 + j = tap j of head h (rows of the
// head's 32 channels), everything else zero.
void pack_final_umma(const FinalW& w, int nheads, std::vector<uint16_t>& out) {
  const int K = nheads * 32;
  out.assign((size_t)(K / 16) * 2 * 16 * 8, 0);
  for (int h = 0; h < nheads; ++h)
    for (int j = 0; j < 7; ++j)
      for (int c = 0; c < 32; ++c) {
        const int k = h * 32 + c;
        const __half hv = __float2half_rn(w.w[h][j][c]);
        uint16_t bits;
        memcpy(&bits, &hv, 2);
        out[((((size_t)(k / 16) * 2 + (k % 16) / 8) * 16) + (h * 8 + j)) * 8 + (k % 8)] = bits;
      }
}

int launch_final_umma(const Act& in, int in_coff8, const __half* w_packed, const FinalW& w, int nheads, float* y, int B, int T,
                      const float* x_lr, cudaStream_t stream) {
  AR_CHECK(nheads == 1 || nheads == 2, AR_ERR_INVALID, "final_umma: one or two heads");
  FinalUmmaArgs a;
  a.in = in.h(); a.in_bs = in.bs; a.in_Tp = in.Tp; a.in_coff8 = in_coff8;
  a.w = w_packed; a.bias[0] = w.bias[0]; a.bias[1] = nheads > 1 ? w.bias[1] : 0.f;
  a.y = y; a.x_lr = x_lr; a.nheads = nheads; a.T = T; a.B = B;
  a.tiles_per_item = (T + FU_OUT - 1) / FU_OUT;
  const int num_tiles = B * a.tiles_per_item;
  const int stage_bytes = nheads * 4 * TILE_M * 16;
  const int smem = 2048 + FU_STAGES * stage_bytes + 2 * TILE_M * FU_SROW * 4 + FU_RING_BYTES + 256;
  static DeviceOnce attrs;
  if (attrs.pending()) {
    AR_CUDA_OK(cudaFuncSetAttribute(final_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attrs.done();
  }
  int grid = sm_count();
  if (grid > num_tiles) grid = num_tiles;
  final_umma_kernel<<<grid, FU_THREADS, smem, stream>>>(a, num_tiles);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
