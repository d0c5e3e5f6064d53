// tcgen05 implicit-GEMM Conv1d engine for sm_100a.
//
//   D[128 time rows, N out channels] (fp32, TMEM) += sum over taps j, input-channel blocks kb
//        A_j,kb[128 x 16] (fp16, smem, K-major, no swizzle)  x  B_j,kb[16 x N] (fp16, smem, K-major)
//
// * Activations live in HBM as H8 ([C/8][Tp][8] fp16, see ar_common.cuh), so the rows a tile needs
//   for ALL taps of one 8-channel chunk are one contiguous run: one cp.async.bulk (TMA, UBLKCP)
//   per chunk.  In shared memory chunk c of a stage sits at c*R*16 bytes (R = 128 + reach), which is
//   exactly the canonical no-swizzle K-major UMMA layout ((8,m),(8,2)) with SBO = 128 B
//   (8 rows x 16 B) and LBO = R*16 B.  Tap j (dilation d) is the same descriptor with its
//   start address advanced by j*d*16 bytes -- no im2col, no per-tap reload.
// * Weights are pre-packed per (channel block, tap) as [2][Ns][4] => LBO = Ns*16 B, SBO = 128 B and
//   stay RESIDENT in shared memory for the whole (persistent) CTA: they are fetched once per CTA,
//   not once per tile, so L2->SM traffic per tile is the activation tile only.  Layers whose
//   weights exceed the budget are split along N into `n_slices` column slices (packed per slice on
//   the host); CTA c owns slice c % n_slices and walks the tiles with stride grid/n_slices, so the
//   CTAs that share an activation tile run side by side and the re-read hits L2.
// * Warp roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer (also zeroes out-of-range
//   rows of edge tiles = conv zero padding), warps 2..5 = epilogue (TMEM -> registers ->
//   bias/LeakyReLU/residual/TF32-round -> coalesced float4 stores, + fused max-pool or
//   2x interleave).  Accumulators are double-buffered in TMEM so the epilogue of tile i
//   overlaps the MMAs of tile i+1.  Persistent grid: one CTA per SM, static tile striding.
#include "ar_common.cuh"
#include "umma_ptx.cuh"
#include "umma_epilogue.cuh"
#include <cstdlib>

namespace ar {

constexpr int EPI_WARPS = 8;                   // two per TMEM lane quarter, each takes half of the columns
constexpr int UMMA_THREADS = 64 + 32 * EPI_WARPS;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int BAR_BYTES = 256;
constexpr int BIAS_BYTES = 1024;              // bias of this CTA's column slice, staged in shared memory

struct UmmaCfg {
  int kbs;          // 8-channel K blocks per pipeline stage
  int stages;
  int R;            // activation rows per chunk per tile
  int a_bytes;      // bytes of one activation stage (= stage_bytes)
  int w_bytes;      // resident weight slice: Cin*taps*Ns*4
  int stage_bytes;
  int ncol;         // TMEM columns per accumulator buffer (pow2 >= Ns)
  int tmem_cols;    // allocated columns (2 buffers)
  int nks;          // pipeline stages per tile = Cin / (8*kbs)
  int smem_bytes;
};

// ----------------------------------------------------------------------------- kernel
// MODE / POOL / RES select the fused epilogue at compile time (ar_common.cuh: ConvMode; max-pool copy;
// residual add); LeakyReLU slope and TF32 rounding stay runtime-uniform.  TAPS is a template parameter
// so the single issuing thread sees a fully unrolled tap loop: the MMAs of a K block go out back to back
// instead of one per ~15 dependent integer instructions.
template <int MODE, bool POOL, bool RES, int TAPS>
__global__ void __launch_bounds__(UMMA_THREADS, 1)
conv_umma_kernel(const __grid_constant__ ConvParams p, const __grid_constant__ UmmaCfg cfg, int num_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // shared memory: [resident weight slice][stages x activation stage][barriers]
  const uint32_t w_base = smem_u32(smem);
  const uint32_t smem_base = w_base + cfg.w_bytes;
  uint8_t* const stage_ptr = smem + cfg.w_bytes;
  const uint32_t bar_base = smem_base + cfg.stages * cfg.stage_bytes;
  // barrier slots (8 bytes each): full[stages], empty[stages], tmem_full[2], tmem_empty[2], weights, tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (cfg.stages + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * cfg.stages + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * cfg.stages + 2 + i); };
  const uint32_t w_bar = bar_base + 8u * (2 * cfg.stages + 4);
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(stage_ptr + cfg.stages * cfg.stage_bytes + 8 * (2 * cfg.stages + 5));
  float* const s_bias = reinterpret_cast<float*>(stage_ptr + cfg.stages * cfg.stage_bytes + BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < cfg.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)cfg.tmem_cols);
  const int Ns = p.N / p.n_slices;                       // GEMM columns of this CTA's weight slice
  const int slice = blockIdx.x % p.n_slices;
  for (int i = threadIdx.x; i < Ns; i += blockDim.x) s_bias[i] = p.bias[slice * Ns + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tpi = p.tiles_per_item;
  const int R = cfg.R;
  const int tile0 = blockIdx.x / p.n_slices;
  const int tile_step = gridDim.x / p.n_slices;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (elect_one()) {
      // the weight slice, once per CTA
      mbar_expect_tx(w_bar, (uint32_t)cfg.w_bytes);
      const char* wsrc = reinterpret_cast<const char*>(p.w) + (size_t)slice * cfg.w_bytes;
      for (int off = 0; off < cfg.w_bytes; off += 32768) {
        const int n = cfg.w_bytes - off < 32768 ? cfg.w_bytes - off : 32768;
        bulk_g2s(w_base + off, wsrc + off, (uint32_t)n, w_bar);
      }
      // per-stage work is kept to: wait, expect_tx, 2*kbs bulk copies, counter bump (no divisions)
      int s = 0;
      uint32_t ph = 0;
      const uint32_t row_bytes = (uint32_t)(R * 16);
      const long long chunk_stride = (long long)p.in_Tp * 8;             // halves between 8-channel chunks
      const int chunks_per_stage = cfg.kbs * 2;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int b = tile / tpi;
        const int t0 = (tile - b * tpi) * TILE_M;
        const __half* src = p.in + act_off(p.in_bs, p.in_Tp, b, p.in_coff8, t0 - p.pad_left);
        for (int ks = 0; ks < cfg.nks; ++ks) {
          const uint32_t fb = full_bar(s);
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(fb, (uint32_t)cfg.stage_bytes);
          uint32_t dst = smem_base + s * cfg.stage_bytes;
          for (int c = 0; c < chunks_per_stage; ++c) {
            bulk_g2s(dst, src, row_bytes, fb);
            dst += row_bytes;
            src += chunk_stride;
          }
          if (++s == cfg.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_f16(TILE_M, Ns);
    const uint64_t a_desc_hi = make_desc(0u, (uint32_t)(R * 16), 128u);
    const uint64_t b_desc_hi = make_desc(0u, (uint32_t)(Ns * 16), 128u);
    mbar_wait(w_bar, 0);
    int s = 0, tl = 0;
    uint32_t ph = 0;
    const uint32_t b_step = (uint32_t)(Ns * 2);            // descriptor units (16 B) between (kb,tap) weight blocks
    const uint32_t a_step = (uint32_t)(2 * R);             // ... between 8-channel K blocks of a stage
    const uint32_t w_addr0 = w_base >> 4;
    const uint32_t dil_u = (uint32_t)p.dil;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tl) {
      const int t0 = (tile % tpi) * TILE_M;
      const int buf = tl & 1;
      const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
      mbar_wait(tempty_bar(buf), aph ^ 1u);
      tc_fence_after();
      const int tfirst = t0 - p.pad_left;                 // time of local row 0
      const bool edge = (tfirst < 0) || (tfirst + R > p.Tin);
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * cfg.ncol);
      uint32_t b_addr = w_addr0;
      uint32_t accum = 0u;
      for (int ks = 0; ks < cfg.nks; ++ks) {
        mbar_wait(full_bar(s), ph);
        if (edge) {  // conv zero padding: rows outside [0, Tin) become zeros (first / last tiles only)
          uint8_t* a_ptr = stage_ptr + s * cfg.stage_bytes;
          for (int r = lane; r < R; r += 32) {
            const int t = tfirst + r;
            if (t < 0 || t >= p.Tin)
              for (int c = 0; c < cfg.kbs * 2; ++c)
                *reinterpret_cast<float4*>(a_ptr + (c * R + r) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          fence_async_smem();
          __syncwarp();
        }
        tc_fence_after();
        if (elect_one()) {
          // descriptors differ only in their 14-bit start-address field (16-byte units)
          uint32_t a_addr = (smem_base + s * cfg.stage_bytes) >> 4;
          for (int kb = 0; kb < cfg.kbs; ++kb) {
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
              umma_f16(d_tmem, a_desc_hi | (uint64_t)(a_addr + (uint32_t)j * dil_u), b_desc_hi | (uint64_t)(b_addr + (uint32_t)j * b_step),
                        idesc, (j == 0) ? accum : 1u);
            }
            accum = 1u;
            b_addr += (uint32_t)TAPS * b_step;
            a_addr += a_step;
          }
          umma_commit(empty_bar(s));                       // frees the smem stage when the MMAs retire
          if (ks == cfg.nks - 1) umma_commit(tfull_bar(buf));  // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++s == cfg.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    // Warp w may touch TMEM lanes [32*(w%4), +32) = GEMM rows; the two warps of a lane quarter split
    // the columns.  Row addresses are hoisted per tile; per 4-column chunk the work is: bias add,
    // LeakyReLU as max(v, slope*v), (residual), (TF32 round), one coalesced float4 store.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;                    // 0 or 1
    const int wcols = Ns >= 32 ? Ns / 2 : Ns;            // columns this warp handles (multiple of 16)
    const int col_lo = Ns >= 32 ? half * wcols : 0;
    const bool active = Ns >= 32 || half == 0;
    const float slope = p.lrelu ? LRELU_SLOPE : 1.0f;
    const int gcol0 = slice * Ns + col_lo;               // first global GEMM column of this warp
    int tl = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++tl) {
      const int b = tile / tpi;
      const int t = (tile % tpi) * TILE_M + q * 32 + lane;
      const int buf = tl & 1;
      const uint32_t aph = (uint32_t)(tl >> 1) & 1u;
      const EpiRow row = epi_row<MODE, POOL, RES>(p, b, t, gcol0);
      uint4 resv[2];
      epi_prefetch_res<RES>(row, active, resv);
      mbar_wait(tfull_bar(buf), aph);
      tc_fence_after();
      if (active) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * cfg.ncol + col_lo);
        epi_store<MODE, POOL, RES>(row, s_bias + col_lo, taddr, wcols, slope, resv);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)cfg.tmem_cols);
}

// ----------------------------------------------------------------------------- host side
static bool pick_cfg(const ConvParams& p, UmmaCfg& c) {
  const int Ns = p.N / p.n_slices;
  c.R = TILE_M + (p.taps - 1) * p.dil;
  int ncol = 32;
  while (ncol < Ns) ncol <<= 1;
  c.ncol = ncol;
  c.tmem_cols = 2 * ncol;
  c.w_bytes = p.Cin * p.taps * Ns * 2;
  const int room = conv_smem_budget() - BAR_BYTES - BIAS_BYTES - c.w_bytes;
  for (int kbs = 4; kbs >= 1; kbs >>= 1) {
    if (p.Cin % (16 * kbs)) continue;
    c.kbs = kbs;
    c.a_bytes = c.stage_bytes = kbs * 2 * c.R * 16;
    int stages = room / c.stage_bytes;
    if (stages > 8) stages = 8;
    if (stages >= 4 || (kbs == 1 && stages >= 2)) {
      c.stages = stages;
      c.nks = p.Cin / (16 * kbs);
      c.smem_bytes = c.w_bytes + stages * c.stage_bytes + BAR_BYTES + BIAS_BYTES;
      return true;
    }
  }
  return false;
}

static int g_conv_smem = 0;
int set_conv_smem_kb(int kb) {
  if (kb < 64 || kb > 227) { set_error("conv shared-memory budget must be within [64, 227] KB"); return AR_ERR_INVALID; }
  g_conv_smem = kb * 1024;
  return AR_OK;
}
int conv_smem_budget() {
  if (g_conv_smem == 0) {
    const char* e = getenv("AR_CONV_SMEM_KB");
    int kb = e ? atoi(e) : 227;
    if (kb < 64) kb = 64;
    if (kb > 227) kb = 227;
    g_conv_smem = kb * 1024;
  }
  return g_conv_smem;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int launch_conv_umma(const ConvParams& p, cudaStream_t stream) {
  AR_CHECK(p.Cin % 16 == 0 && p.n_slices >= 1 && p.N % (16 * p.n_slices) == 0 && p.N >= 16 && p.N <= 256, AR_ERR_INVALID,
           "conv_umma: unsupported channel counts");
  AR_CHECK(p.res == nullptr || p.N / p.n_slices <= 32, AR_ERR_INVALID, "conv_umma: residual epilogue supports at most 32 columns per slice");
  AR_CHECK(p.pad_left <= HALO && (p.taps - 1) * p.dil - p.pad_left <= HALO, AR_ERR_INVALID, "conv_umma: conv reach exceeds HALO");
  UmmaCfg cfg;
  AR_CHECK(pick_cfg(p, cfg), AR_ERR_INVALID, "conv_umma: no pipeline configuration fits shared memory");
  AR_CHECK(p.mode == MODE_SAME || (p.pool == nullptr && p.res == nullptr), AR_ERR_INVALID, "conv_umma: interleave mode has no pool/residual epilogue");
  AR_CHECK(p.mode == MODE_SAME || (p.N / 2) % (p.N / p.n_slices >= 32 ? p.N / p.n_slices / 2 : p.N / p.n_slices) == 0, AR_ERR_INVALID,
           "conv_umma: interleave phases must align with the epilogue column split");
  AR_CHECK(!(p.pool && p.res), AR_ERR_INVALID, "conv_umma: pool and residual epilogues are exclusive");
  const int num_tiles = p.B * p.tiles_per_item;
  int groups = sm_count() / p.n_slices;                 // CTAs per slice
  if (groups > num_tiles) groups = num_tiles;
  const int grid = groups * p.n_slices;
  using Kernel = void (*)(ConvParams, UmmaCfg, int);
  struct Entry { int variant, taps; Kernel k; };
  static const Entry table[] = {
      {EV_PLAIN, 1, conv_umma_kernel<MODE_SAME, false, false, 1>}, {EV_PLAIN, 3, conv_umma_kernel<MODE_SAME, false, false, 3>},
      {EV_PLAIN, 5, conv_umma_kernel<MODE_SAME, false, false, 5>}, {EV_PLAIN, 7, conv_umma_kernel<MODE_SAME, false, false, 7>},
      {EV_POOL, 3, conv_umma_kernel<MODE_SAME, true, false, 3>},   {EV_RES, 3, conv_umma_kernel<MODE_SAME, false, true, 3>},
      {EV_INTERLEAVE, 1, conv_umma_kernel<MODE_INTERLEAVE2, false, false, 1>},
      {EV_INTERLEAVE, 3, conv_umma_kernel<MODE_INTERLEAVE2, false, false, 3>},
  };
  static bool attr_set = false;
  if (!attr_set) {
    for (const Entry& e : table) AR_CUDA_OK(cudaFuncSetAttribute(e.k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
    attr_set = true;
  }
  const int variant = epi_variant(p);
  Kernel kernel = nullptr;
  for (const Entry& e : table)
    if (e.variant == variant && e.taps == p.taps) kernel = e.k;
  AR_CHECK(kernel != nullptr, AR_ERR_INVALID, "conv_umma: no kernel instantiated for this (epilogue, taps) combination");
  kernel<<<grid, UMMA_THREADS, cfg.smem_bytes, stream>>>(p, cfg, num_tiles);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
