// Fused conv chains on the 2-CTA tcgen05 engine: a k-tap (dilated) Conv1d followed by one or two pointwise
// (k = 1) convs -- the StereoSeparator's `_dilated_block` (stereo_separator.py:49-64: conv k3 dilated + BN +
// LeakyReLU, conv k1 + BN + LeakyReLU) and, for the last block, the LSTM input projection behind it
// (stereo_separator.py:104-106) -- computed per 256-row tile pair WITHOUT writing the intermediate
// activations to HBM:
//
//   G1: D1[256 x N1] (TMEM) = sum_taps A(rows from HBM via bulk copies, tap-shifted descriptors) * W1
//   E1: D1 -> +bias, LeakyReLU, fp16 -> shared memory, laid out as the K-major no-swizzle A operand of G2
//   G2: D2[256 x N2] = I1 * W2            (A straight from shared memory, W2 resident like W1)
//   E2: last stage: the usual fused epilogue to HBM (H8 or time-blocked);  otherwise -> I2 and G3 / E3.
//
// The unfused layers of these blocks are HBM-bound (k1 128->128: 512 B of activation traffic per row for
// 32 K MACs), so removing the intermediate write + read halves the time of a block; the three-GEMM chain
// (128 -k3-> 128 -k1-> 128 -k1-> 256) moves 768 B per row instead of 1 792.
//
// Same roles as conv_umma2.cu: warp 0 = bulk-copy producer, warp 1 = MMA warp (leader CTA issues; the peer
// zero-pads its edge rows and forwards "my half is ready" arrivals), warps 2..9 = epilogues.  The GEMMs of
// consecutive tile pairs are software-pipelined so the tensor pipe does not idle while an epilogue turns an
// accumulator into the next operand:
//   two GEMMs  : slot s issues G1(s), G2(s-1); accumulators and I1 double-buffered
//   three GEMMs: slot s issues G3(s-2), G2(s-1), G1(s); every buffer single (TMEM: 128+128+256 = 512 columns) --
//                in that order each epilogue has a full GEMM of another tile to hide behind.
#include "ar_common.cuh"
#include "umma_ptx.cuh"
#include "umma_epilogue.cuh"

namespace ar {

constexpr int CH_EPI_WARPS = 8;
constexpr int CH_THREADS = 64 + 32 * CH_EPI_WARPS;
constexpr int CH_SMEM_BUDGET = 227 * 1024;
constexpr int CH_BAR_BYTES = 512;
constexpr int CH_BIAS_BYTES = 2048;      // <= 512 fp32 biases over all stages
constexpr int CH_RI = TILE_M;            // rows of an intermediate operand (pointwise follow-up convs: no halo)

struct ChainCfg {
  int kbs, stages, R, nks, stage_bytes;  // activation ring of the first GEMM
  int w_bytes[3], w_off[3];              // resident weight halves per CTA
  int i_off[2], i_bytes[2], nbI[2];      // intermediate operands (output of GEMM g = input of GEMM g+1)
  int acc_col[3], nbA[3];                // TMEM columns: GEMM g buffer b at acc_col[g] + b * N[g]
  int tmem_cols;
  int stage_off, bar_off, bias_off, smem_bytes;
};

// buffer index / mbarrier phase of the it-th use of a set of nb (1 or 2) buffers
__device__ __forceinline__ int buf_of(int it, int nb) { return it & (nb - 1); }
__device__ __forceinline__ uint32_t phase_of(int it, int nb) { return (uint32_t)(it >> (nb - 1)) & 1u; }

template <int TAPS, int NG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CH_THREADS, 1)
conv_chain_kernel(const __grid_constant__ ChainParams cp, const __grid_constant__ ChainCfg cfg, int num_pairs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvParams& p = cp.p;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t stage_base = sbase + cfg.stage_off;
  uint8_t* const stage_ptr = smem + cfg.stage_off;
  const uint32_t bar_base = sbase + cfg.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto peer_bar = [&](int s) { return bar_base + 8u * (16 + s); };            // leader only
  auto tfull_bar = [&](int g, int b) { return bar_base + 8u * (24 + g * 2 + b); };
  auto tempty_bar = [&](int g, int b) { return bar_base + 8u * (30 + g * 2 + b); };   // leader only
  auto ifull_bar = [&](int g, int b) { return bar_base + 8u * (36 + g * 2 + b); };    // own epilogue warps -> own MMA warp
  auto ipeer_bar = [&](int g, int b) { return bar_base + 8u * (40 + g * 2 + b); };    // peer's MMA warp -> leader
  auto iempty_bar = [&](int g, int b) { return bar_base + 8u * (44 + g * 2 + b); };
  const uint32_t w_bar = bar_base + 8u * 48;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + cfg.bar_off + 8 * 49);
  float* const s_bias = reinterpret_cast<float*>(smem + cfg.bias_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < cfg.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(peer_bar(s), 1);
    }
    for (int g = 0; g < NG; ++g)
      for (int b = 0; b < 2; ++b) {
        mbar_init(tfull_bar(g, b), 1);
        mbar_init(tempty_bar(g, b), 2 * CH_EPI_WARPS);
      }
    for (int g = 0; g < NG - 1; ++g)
      for (int b = 0; b < 2; ++b) {
        mbar_init(ifull_bar(g, b), CH_EPI_WARPS);
        mbar_init(ipeer_bar(g, b), 1);
        mbar_init(iempty_bar(g, b), 1);
      }
    mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc2(smem_u32((const void*)tmem_slot), (uint32_t)cfg.tmem_cols);
  {
    int off = 0;
    for (int g = 0; g < NG; ++g) {
      for (int i = threadIdx.x; i < cp.N[g]; i += blockDim.x) s_bias[off + i] = cp.bias[g][i];
      off += cp.N[g];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tpi = p.tiles_per_item;
  const int ppi = (tpi + 1) >> 1;
  const int R = cfg.R;
  const int pair0 = blockIdx.x >> 1;
  const int pair_step = gridDim.x >> 1;
  const int n_local = pair0 < num_pairs ? (num_pairs - pair0 + pair_step - 1) / pair_step : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: resident weights, then the activation ring
    if (elect_one()) {
      uint32_t total = 0;
      for (int g = 0; g < NG; ++g) total += (uint32_t)cfg.w_bytes[g];
      mbar_expect_tx(w_bar, total);
      for (int g = 0; g < NG; ++g) {
        const char* wsrc = reinterpret_cast<const char*>(cp.w[g]) + (size_t)rank * cfg.w_bytes[g];
        for (int off = 0; off < cfg.w_bytes[g]; off += 32768) {
          const int n = cfg.w_bytes[g] - off < 32768 ? cfg.w_bytes[g] - off : 32768;
          bulk_g2s(sbase + cfg.w_off[g] + off, wsrc + off, (uint32_t)n, w_bar);
        }
      }
      int s = 0;
      uint32_t ph = 0;
      const uint32_t row_bytes = (uint32_t)(R * 16);
      const long long chunk_stride = (long long)p.in_Tp * 8;
      const int chunks_per_stage = cfg.kbs * 2;
      for (int it = 0; it < n_local; ++it) {
        const int pr = pair0 + it * pair_step;
        const int b = pr / ppi;
        int tl_in_item = (pr - b * ppi) * 2 + (int)rank;
        if (tl_in_item > tpi - 1) tl_in_item = tpi - 1;   // odd tile count: the idle half re-reads a valid tile (rows get zeroed)
        const int t0 = tl_in_item * TILE_M;
        const __half* src = p.in + act_off(p.in_bs, p.in_Tp, b, p.in_coff8, t0 - p.pad_left);
        for (int ks = 0; ks < cfg.nks; ++ks) {
          const uint32_t fb = full_bar(s);
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(fb, (uint32_t)cfg.stage_bytes);
          uint32_t dst = stage_base + s * cfg.stage_bytes;
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
    // ------------------------------------------------------------------ MMA warp
    mbar_wait(w_bar, 0);
    int s = 0;
    uint32_t ph = 0;
    // first GEMM of tile pair `it`: k-tap conv, rows streamed through the ring
    auto gemm_first = [&](int it) {
      const int N1 = cp.N[0], Nh = N1 >> 1;
      const uint32_t idesc = make_idesc_f16(256, N1);
      const uint64_t a_desc_hi = make_desc(0u, (uint32_t)(R * 16), 128u);
      const uint64_t b_desc_hi = make_desc(0u, (uint32_t)(Nh * 16), 128u);
      const uint32_t b_step = (uint32_t)(Nh * 2);
      const uint32_t a_step = (uint32_t)(2 * R);
      const uint32_t dil_u = (uint32_t)p.dil;
      const int pr = pair0 + it * pair_step;
      const int b = pr / ppi;
      const int tl_in_item = (pr - b * ppi) * 2 + (int)rank;
      const int t0 = (tl_in_item > tpi - 1 ? tpi - 1 : tl_in_item) * TILE_M;
      const int tfirst = t0 - p.pad_left;
      const bool dead = tl_in_item > tpi - 1;
      const bool edge = dead || (tfirst < 0) || (tfirst + R > p.Tin);
      const int buf = buf_of(it, cfg.nbA[0]);
      if (leader) {
        mbar_wait(tempty_bar(0, buf), phase_of(it, cfg.nbA[0]) ^ 1u);
        tc_fence_after();
      }
      const uint32_t d_tmem = tmem_base + (uint32_t)(cfg.acc_col[0] + buf * N1);
      uint32_t b_addr = (sbase + cfg.w_off[0]) >> 4;
      uint32_t accum = 0u;
      for (int ks = 0; ks < cfg.nks; ++ks) {
        mbar_wait(full_bar(s), ph);
        if (edge) {  // conv zero padding of this CTA's rows
          uint8_t* a_ptr = stage_ptr + s * cfg.stage_bytes;
          for (int r = lane; r < R; r += 32) {
            const int t = tfirst + r;
            if (dead || t < 0 || t >= p.Tin)
              for (int c = 0; c < cfg.kbs * 2; ++c)
                *reinterpret_cast<float4*>(a_ptr + (c * R + r) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          fence_async_smem();
          __syncwarp();
        }
        if (!leader) {
          if (elect_one()) {
            if (edge) mbar_arrive_remote_release(mapa_u32(peer_bar(s), 0));
            else mbar_arrive_remote(mapa_u32(peer_bar(s), 0));
          }
          __syncwarp();
        } else {
          mbar_wait(peer_bar(s), ph);
          tc_fence_after();
          if (elect_one()) {
            uint32_t a_addr = (stage_base + s * cfg.stage_bytes) >> 4;
            for (int kb = 0; kb < cfg.kbs; ++kb) {
#pragma unroll
              for (int j = 0; j < TAPS; ++j)
                umma2_f16(d_tmem, a_desc_hi | (uint64_t)(a_addr + (uint32_t)j * dil_u),
                          b_desc_hi | (uint64_t)(b_addr + (uint32_t)j * b_step), idesc, (j == 0) ? accum : 1u);
              accum = 1u;
              b_addr += (uint32_t)TAPS * b_step;
              a_addr += a_step;
            }
            umma_commit2(empty_bar(s));
            if (ks == cfg.nks - 1) umma_commit2(tfull_bar(0, buf));
          }
          __syncwarp();
        }
        if (++s == cfg.stages) { s = 0; ph ^= 1u; }
      }
    };
    // pointwise GEMM g (1 or 2) of tile pair `it`: A = intermediate operand written by the epilogue of GEMM g-1
    auto gemm_next = [&](int g, int it) {
      const int K = cp.N[g - 1], Ng = cp.N[g], Nh = Ng >> 1;
      const int bi = buf_of(it, cfg.nbI[g - 1]);
      const uint32_t iph = phase_of(it, cfg.nbI[g - 1]);
      mbar_wait(ifull_bar(g - 1, bi), iph);                 // this CTA's 128 operand rows are in shared memory
      if (!leader) {
        if (elect_one()) mbar_arrive_remote_release(mapa_u32(ipeer_bar(g - 1, bi), 0));
        __syncwarp();
        return;
      }
      mbar_wait_cluster(ipeer_bar(g - 1, bi), iph);
      const int buf = buf_of(it, cfg.nbA[g]);
      mbar_wait(tempty_bar(g, buf), phase_of(it, cfg.nbA[g]) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t idesc = make_idesc_f16(256, Ng);
        const uint64_t a_desc_hi = make_desc(0u, (uint32_t)(CH_RI * 16), 128u);
        const uint64_t b_desc_hi = make_desc(0u, (uint32_t)(Nh * 16), 128u);
        const uint32_t d_tmem = tmem_base + (uint32_t)(cfg.acc_col[g] + buf * Ng);
        uint32_t a_addr = (sbase + cfg.i_off[g - 1] + bi * cfg.i_bytes[g - 1]) >> 4;
        uint32_t b_addr = (sbase + cfg.w_off[g]) >> 4;
        for (int kb = 0; kb < K / 16; ++kb) {
          umma2_f16(d_tmem, a_desc_hi | (uint64_t)a_addr, b_desc_hi | (uint64_t)b_addr, idesc, kb ? 1u : 0u);
          a_addr += (uint32_t)(2 * CH_RI);
          b_addr += (uint32_t)(Nh * 2);
        }
        umma_commit2(iempty_bar(g - 1, bi));                // both CTAs may overwrite this operand buffer
        umma_commit2(tfull_bar(g, buf));
      }
      __syncwarp();
    };
    for (int slot = 0; slot < n_local + NG - 1; ++slot) {
      if (NG == 2) {
        if (slot < n_local) gemm_first(slot);
        if (slot >= 1) gemm_next(1, slot - 1);
      } else {
        if (slot >= 2) gemm_next(2, slot - 2);
        if (slot >= 1 && slot - 1 < n_local) gemm_next(1, slot - 1);
        if (slot < n_local) gemm_first(slot);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps: own 128 rows
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    // accumulator of GEMM g -> bias, LeakyReLU, fp16 -> this CTA's operand rows of GEMM g+1
    auto epi_mid = [&](int g, int it, int bias_off) {
      const int Ng = cp.N[g];
      const int wcols = Ng >> 1, col_lo = half * wcols;      // Ng >= 32
      const float slope = cp.lrelu[g] ? LRELU_SLOPE : 1.0f;
      const int buf = buf_of(it, cfg.nbA[g]);
      const int bi = buf_of(it, cfg.nbI[g]);
      mbar_wait(tfull_bar(g, buf), phase_of(it, cfg.nbA[g]));
      mbar_wait(iempty_bar(g, bi), phase_of(it, cfg.nbI[g]) ^ 1u);   // GEMM g+1 of the previous user has read the buffer
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cfg.acc_col[g] + buf * Ng + col_lo);
      uint8_t* const dst = smem + cfg.i_off[g] + bi * cfg.i_bytes[g] + (q * 32 + lane) * 16;
      const float* bw = s_bias + bias_off + col_lo;
      for (int cb = 0; cb < wcols; cb += 32) {
        uint32_t a[32];
        const int ncol = wcols - cb < 32 ? 16 : 32;
        if (ncol == 32) tmem_ld32_nowait(taddr + cb, a);
        else tmem_ld16_nowait(taddr + cb, a);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (8 * c < ncol) {
            const float4 b0 = *reinterpret_cast<const float4*>(bw + cb + 8 * c);
            const float4 b1 = *reinterpret_cast<const float4*>(bw + cb + 8 * c + 4);
            float v[8] = {__uint_as_float(a[8 * c]) + b0.x,     __uint_as_float(a[8 * c + 1]) + b0.y,
                          __uint_as_float(a[8 * c + 2]) + b0.z, __uint_as_float(a[8 * c + 3]) + b0.w,
                          __uint_as_float(a[8 * c + 4]) + b1.x, __uint_as_float(a[8 * c + 5]) + b1.y,
                          __uint_as_float(a[8 * c + 6]) + b1.z, __uint_as_float(a[8 * c + 7]) + b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], slope * v[i]);
            const int chunk = ((col_lo + cb) >> 3) + c;
            *reinterpret_cast<uint4*>(dst + chunk * (CH_RI * 16)) = pack_half8(v);
          }
        }
      }
      tc_fence_before();
      fence_async_smem();                                    // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_remote(mapa_u32(tempty_bar(g, buf), 0));
        mbar_arrive(ifull_bar(g, bi));
      }
    };
    // accumulator of the last GEMM -> the usual fused epilogue to HBM
    auto epi_last = [&](int it, int bias_off) {
      const int g = NG - 1;
      const int Ng = cp.N[g];
      const int wcols = Ng >> 1, col_lo = half * wcols;
      const float slope = cp.lrelu[g] ? LRELU_SLOPE : 1.0f;
      const int pr = pair0 + it * pair_step;
      const int b = pr / ppi;
      const int tl_in_item = (pr - b * ppi) * 2 + (int)rank;
      const int t = tl_in_item * TILE_M + q * 32 + lane;     // >= Tin for a dead tile => every store is masked
      const EpiRow row = epi_row<MODE_SAME, false, false>(cp.pl, b, t, col_lo);
      const int buf = buf_of(it, cfg.nbA[g]);
      uint4 resv[2];
      mbar_wait(tfull_bar(g, buf), phase_of(it, cfg.nbA[g]));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cfg.acc_col[g] + buf * Ng + col_lo);
      epi_store<MODE_SAME, false, false>(row, s_bias + bias_off + col_lo, taddr, wcols, slope, resv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_u32(tempty_bar(g, buf), 0));
    };
    const int boff1 = cp.N[0], boff2 = cp.N[0] + cp.N[1];
    for (int slot = 0; slot < n_local + NG - 1; ++slot) {
      if (NG == 2) {
        if (slot < n_local) epi_mid(0, slot, 0);
        if (slot >= 1) epi_last(slot - 1, boff1);
      } else {
        if (slot >= 2) epi_last(slot - 2, boff2);
        if (slot >= 1 && slot - 1 < n_local) epi_mid(1, slot - 1, boff1);
        if (slot < n_local) epi_mid(0, slot, 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc2(tmem_base, (uint32_t)cfg.tmem_cols);
}

// ----------------------------------------------------------------------------- host side
static bool pick_chain_cfg(const ChainParams& cp, ChainCfg& c) {
  const ConvParams& p = cp.p;
  const int NG = cp.n_gemms;
  c.R = TILE_M + (p.taps - 1) * p.dil;
  int off = 0, cols = 0;
  for (int g = 0; g < NG; ++g) {
    const int K = g == 0 ? p.Cin * p.taps : cp.N[g - 1];
    c.w_bytes[g] = K * (cp.N[g] / 2) * 2;
    c.w_off[g] = off;
    off += (c.w_bytes[g] + 1023) / 1024 * 1024;
    c.nbA[g] = NG == 2 ? 2 : 1;
    c.acc_col[g] = cols;
    cols += c.nbA[g] * cp.N[g];
  }
  if (cols > 512) return false;
  int tc = 32;
  while (tc < cols) tc <<= 1;
  c.tmem_cols = tc;
  for (int g = 0; g < NG - 1; ++g) {
    c.nbI[g] = NG == 2 ? 2 : 1;
    c.i_bytes[g] = (cp.N[g] / 8) * CH_RI * 16;
    c.i_off[g] = off;
    off += c.nbI[g] * c.i_bytes[g];
  }
  c.stage_off = off;
  const int room = CH_SMEM_BUDGET - CH_BAR_BYTES - CH_BIAS_BYTES - off;
  for (int kbs = 4; kbs >= 1; kbs >>= 1) {
    if (p.Cin % (16 * kbs)) continue;
    c.kbs = kbs;
    c.stage_bytes = kbs * 2 * c.R * 16;
    int stages = room / c.stage_bytes;
    if (stages > 8) stages = 8;
    if (stages >= 4 || (kbs == 1 && stages >= 2)) {
      c.stages = stages;
      c.nks = p.Cin / (16 * kbs);
      c.bar_off = c.stage_off + stages * c.stage_bytes;
      c.bar_off = (c.bar_off + 15) / 16 * 16;
      c.bias_off = c.bar_off + CH_BAR_BYTES;
      c.smem_bytes = c.bias_off + CH_BIAS_BYTES;
      return true;
    }
  }
  return false;
}

int launch_conv_chain(const ChainParams& cp, cudaStream_t stream) {
  const ConvParams& p = cp.p;
  const int NG = cp.n_gemms;
  AR_CHECK(NG == 2 || NG == 3, AR_ERR_INVALID, "conv_chain: 2 or 3 GEMMs");
  AR_CHECK(p.Cin % 16 == 0 && p.mode == MODE_SAME && p.pool == nullptr && p.res == nullptr, AR_ERR_INVALID, "conv_chain: unsupported first layer");
  AR_CHECK(p.pad_left <= HALO && (p.taps - 1) * p.dil - p.pad_left <= HALO, AR_ERR_INVALID, "conv_chain: conv reach exceeds HALO");
  int nb = 0;
  for (int g = 0; g < NG; ++g) {
    AR_CHECK(cp.N[g] % 32 == 0 && cp.N[g] >= 32 && cp.N[g] <= 256, AR_ERR_INVALID, "conv_chain: unsupported channel count");
    AR_CHECK(g == NG - 1 || cp.N[g] <= 128, AR_ERR_INVALID, "conv_chain: intermediate wider than 128 channels");
    nb += cp.N[g];
  }
  AR_CHECK(nb * 4 <= CH_BIAS_BYTES, AR_ERR_INVALID, "conv_chain: too many bias entries");
  AR_CHECK(cp.pl.res == nullptr || false, AR_ERR_INVALID, "conv_chain: no residual epilogue");
  ChainCfg cfg;
  AR_CHECK(pick_chain_cfg(cp, cfg), AR_ERR_INVALID, "conv_chain: no configuration fits shared memory / TMEM");
  using Kernel = void (*)(ChainParams, ChainCfg, int);
  struct Entry { int taps, ng; Kernel k; };
  static const Entry table[] = {
      {3, 2, conv_chain_kernel<3, 2>}, {3, 3, conv_chain_kernel<3, 3>}, {1, 2, conv_chain_kernel<1, 2>},
  };
  static bool attr_set = false;
  if (!attr_set) {
    for (const Entry& e : table) AR_CUDA_OK(cudaFuncSetAttribute(e.k, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BUDGET));
    attr_set = true;
  }
  Kernel kernel = nullptr;
  for (const Entry& e : table)
    if (e.taps == p.taps && e.ng == NG) kernel = e.k;
  AR_CHECK(kernel != nullptr, AR_ERR_INVALID, "conv_chain: no kernel instantiated for this (taps, stages) combination");
  const int ppi = (p.tiles_per_item + 1) / 2;
  const int num_pairs = p.B * ppi;
  int groups = sm_count() / 2;
  if (groups > num_pairs) groups = num_pairs;
  if (groups < 1) groups = 1;
  kernel<<<groups * 2, CH_THREADS, cfg.smem_bytes, stream>>>(cp, cfg, num_pairs);
  AR_CUDA_OK(cudaGetLastError());
  return AR_OK;
}

}  // namespace ar
