// Weight (and bias) gradient on CTA PAIRS: tcgen05.mma.cta_group::2, M = 256 wide channels over two SMs.
//
//   dW[tap][m][n] = sum_pixels  Wide[p][m] * Narrow[p + off(tap)][n]            (see wgrad_tc.cu for the formulation)
//
// wgrad_tc.cu gives every CTA one 128-channel tile of the wide operand, ALL narrow channels and a group of taps; ncu
// showed it bound by what each SM has to pull in from L2 per 64-pixel step (r1c_wg_l1c2: 13.5x the DRAM bytes cross
// the L2 -> SM crossbar at its 7 TB/s ceiling) and by shared-memory operand reads of its small-N MMAs.  Here
//   * two CTAs of a cluster issue ONE M = 256 MMA per K step: each loads only its own 128 wide channels and HALF of
//     the narrow channels of the pair's N slice (B is split along N over the pair), the accumulator rows of each CTA
//     stay in its own TMEM;
//   * the narrow operand is cut in N slices of <= 128 channels (<= 64 per CTA = one swizzle box) and a pair keeps as
//     many taps of its slice as fit in 512 TMEM columns -- "N slices x taps" instead of "all N x few taps" halves
//     the bytes every pixel costs (256x192x9: 6.6 KB -> 3.2 KB per pixel summed over the CTAs that read it);
//   * the halo'd narrow box is 10 pixels wide instead of 16 (the swizzle is a function of the shared-memory address
//     bits, so any 128-byte row pitch works as the K-group stride): 100 instead of 160 pixels per 64-pixel step;
//   * the bias gradient (column sums of dy) is one more accumulator: an MMA against a constant tile of ones, so the
//     epilogue warps never touch the operand ring and the peer CTA needs no "stage landed" barrier of its own.
// Partials [split][tap][Cw][Cn] and the fixed-order reduction are those of wgrad_tc.cu (bit-deterministic).
//
// Replaces the weight/bias gradients autograd derives for nn.Conv2d in subnet_conv / subnet_conv_1x1
// (/root/reference/archs.py:11-17) and DenseBlock (/root/reference/archs.py:77-81).
#include <stdlib.h>
#include "tc_common.cuh"

namespace sininn {
namespace tc {

constexpr int WP_THREADS = 64 + 8 * 32;                           // TMA warp, MMA warp, 8 epilogue warps
constexpr int WP_BLK = 8;                                         // 8x8 pixel block = one K step of 64 pixels
constexpr uint32_t WP_WIDE_BYTES = 2 * WP_BLK * WP_BLK * 128;     // two 64-channel boxes of 64 pixels
constexpr uint32_t WP_ONES_BYTES = 4096;
constexpr int WP_MAX_GROUPS = 9;

struct WgPairParams {
  int B, H, W;
  int taps, narrow_is_x;
  int Cw, Cn;
  int n_pair, n_half, n_slices;  // narrow channels per pair and tap / per CTA; slices of the narrow operand
  int tap_groups, m_pairs, splits;
  int tap_begin[WP_MAX_GROUPS + 1];
  int blocks_h, blocks_w;
  long long num_blocks, blocks_per_split;
  int stages, halo_w;            // halo_w: pixels per row of the narrow box (3x3: 10 or 16; 1x1: 8)
  uint32_t stage_bytes, narrow_bytes, tx_bytes;
  float* partial;                // [split][tap][Cw][Cn]
  int bias_mode;                 // 0 off, 1 dy is the wide operand, 2 dy is the narrow operand
  float* bias_partial;           // [split][Cout]
  int tma_out;                   // 1: partial tiles leave through TMA tensor stores (needs Cn % 4 == 0)
  long long* trace;              // debugging aid (sininn_debug_set_trace): clock64 stamps of pair 0's roles, or NULL
};

static long long* g_wg_trace = nullptr;
void set_wgrad_pair_trace(long long* buf) { g_wg_trace = buf; }
__device__ __forceinline__ void wp_stamp(const WgPairParams& p, int slot, int lane) {
  if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[slot] = clock64();
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

struct __align__(8) WgPairBarriers {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base, pad;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WP_THREADS, 1)
wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmN,
                  const __grid_constant__ CUtensorMap tmP, const WgPairParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* ring = smem_raw + pad;
  uint8_t* ones = ring + (size_t)p.stages * p.stage_bytes;                    // 1024-aligned (stage_bytes % 1024 == 0)
  WgPairBarriers* bars = reinterpret_cast<WgPairBarriers*>(ones + WP_ONES_BYTES);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t ring_u32 = smem_u32(ring);
  const int rank = (int)cluster_ctarank();

  // work item of this pair
  int q = blockIdx.x >> 1;
  const int split = q % p.splits; q /= p.splits;
  const int tg = q % p.tap_groups; q /= p.tap_groups;
  const int ns = q % p.n_slices;
  const int mp = q / p.n_slices;
  const int tap0 = p.tap_begin[tg];
  const int ntap = p.tap_begin[tg + 1] - tap0;
  const int c0 = ns * p.n_pair;                                               // first narrow channel of the slice
  const long long blk0 = (long long)split * p.blocks_per_split;
  long long blk1 = blk0 + p.blocks_per_split;
  if (blk1 > p.num_blocks) blk1 = p.num_blocks;
  const long long nblk = blk1 > blk0 ? blk1 - blk0 : 0;
  const bool do_bias = p.bias_mode != 0 && tg == 0 && (p.bias_mode == 1 ? ns == 0 : mp == 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmN);
    if (p.tma_out) tma_prefetch_desc(&tmP);
  }
  wp_stamp(p, 0, threadIdx.x);
  // constant tile of ones (bf16 1.0 = 0x3F80): the other operand of the bias-gradient MMAs; a tile that is all
  // ones looks the same under every swizzle
  for (int i = threadIdx.x; i < (int)(WP_ONES_BYTES / 16); i += WP_THREADS)
    reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  wp_stamp(p, 1, threadIdx.x);
  pdl_wait();
  pdl_trigger();
  wp_stamp(p, 2, threadIdx.x);

  if (warp == 0) {
    // ======================= TMA producer (both CTAs): own wide channels, own half of the narrow slice ==========
    int stage = 0; uint32_t phase = 0;
    const int ho = (p.taps == 9) ? 1 : 0;
    const int wch = mp * 256 + rank * 128;
    const int nch = c0 + rank * p.n_half;
    for (long long i = 0; i < nblk; ++i) {
      long long blk = blk0 + i;
      const int bw = (int)(blk % p.blocks_w); blk /= p.blocks_w;
      const int bh = (int)(blk % p.blocks_h);
      const int b = (int)(blk / p.blocks_h);
      const int w0 = bw * WP_BLK, h0 = bh * WP_BLK;
      mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(smem_u32(&bars->full[stage]), 2u * p.tx_bytes);
        const uint32_t full_l = mapa_u32(smem_u32(&bars->full[stage]), 0);     // the leader's barrier
        const uint32_t dst = ring_u32 + stage * p.stage_bytes;
        tma_load_4d_2sm(dst, &tmW, full_l, wch, w0, h0, b);
        tma_load_4d_2sm(dst + WP_WIDE_BYTES / 2, &tmW, full_l, wch + 64, w0, h0, b);
        tma_load_4d_2sm(dst + WP_WIDE_BYTES, &tmN, full_l, nch, w0 - ho, h0 - ho, b);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    wp_stamp(p, 3, lane);
  } else if (warp == 1) {
    if (rank == 0) {
      // ======================= MMA issuer (leader; whole warp loops, predicated issue) =======================
      // D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 256 over the pair
      const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t idesc = idesc_base | ((uint32_t)(p.n_pair >> 3) << 17);
      const uint32_t idesc_b1 = idesc_base | ((uint32_t)(16 >> 3) << 17);        // bias, dy wide: N = 16 columns of ones
      const bool halo = p.taps == 9;
      // MN-major SWIZZLE_128B: LBO = distance between 64-channel boxes, SBO = distance between 8-pixel K groups
      uint64_t a_desc0 = make_desc(ring_u32, 1024, 2);
      a_desc0 = (a_desc0 & ~((uint64_t)0x3FFF << 16)) | ((uint64_t)((WP_WIDE_BYTES / 2) >> 4) << 16);
      uint64_t b_desc0 = make_desc(ring_u32 + WP_WIDE_BYTES, halo ? (uint32_t)p.halo_w * 128u : 1024u, 2);
      uint64_t one_a = make_desc(smem_u32(ones), 1024, 2);                        // 128 "channels" x 16 pixels of ones
      one_a = (one_a & ~((uint64_t)0x3FFF << 16)) | ((uint64_t)(2048 >> 4) << 16);
      const uint64_t one_b = make_desc(smem_u32(ones), 1024, 2);
      const uint32_t b_kstep = (halo ? 2u * (uint32_t)p.halo_w * 128u : 2048u) >> 4;   // 16 pixels further along K
      uint32_t tap_off[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        int sy = 0, sx = 0;
        if (halo) {
          const int tap = tap0 + t;
          sy = tap / 3 - 1; sx = tap % 3 - 1;
          if (!p.narrow_is_x) { sy = -sy; sx = -sx; }
          sy += 1; sx += 1;
        }
        tap_off[t] = (uint32_t)(sy * p.halo_w + sx) * 8u;
      }
      const uint32_t center_off = halo ? (uint32_t)(p.halo_w + 1) * 8u : 0u;
      const uint32_t stage_step = p.stage_bytes >> 4;
      const uint32_t d_bias = tmem_base + (uint32_t)(ntap * p.n_pair);
      const int bias_mode = do_bias ? p.bias_mode : 0;
      int stage = 0; uint32_t phase = 0;
      for (long long i = 0; i < nblk; ++i) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        tc_fence_after();
        if (i == 0) wp_stamp(p, 4, lane);
        __syncwarp();
        const uint32_t lead = elect_pred();
        const uint64_t ad = a_desc0 + (uint64_t)(stage * stage_step);
        const uint64_t bs = b_desc0 + (uint64_t)(stage * stage_step);
        const uint32_t acc = i != 0 ? 1u : 0u;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          if (t < ntap) {
            const uint64_t bd = bs + tap_off[t];
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * p.n_pair);
            umma_bf16_2sm_p(lead, d_tmem, ad, bd, idesc, acc);
            umma_bf16_2sm_p(lead, d_tmem, ad + 128, bd + b_kstep, idesc, 1u);
            umma_bf16_2sm_p(lead, d_tmem, ad + 256, bd + 2 * b_kstep, idesc, 1u);
            umma_bf16_2sm_p(lead, d_tmem, ad + 384, bd + 3 * b_kstep, idesc, 1u);
          }
        }
        if (bias_mode == 1) {                 // colsum of the wide operand: D[m][0..15] += sum_k Wide[k][m] * 1
          umma_bf16_2sm_p(lead, d_bias, ad, one_b, idesc_b1, acc);
          umma_bf16_2sm_p(lead, d_bias, ad + 128, one_b, idesc_b1, 1u);
          umma_bf16_2sm_p(lead, d_bias, ad + 256, one_b, idesc_b1, 1u);
          umma_bf16_2sm_p(lead, d_bias, ad + 384, one_b, idesc_b1, 1u);
        } else if (bias_mode == 2) {          // colsum of the narrow operand (unshifted): D[*][n] += sum_k 1 * Narrow[k][n]
          const uint64_t bd = bs + center_off;
          umma_bf16_2sm_p(lead, d_bias, one_a, bd, idesc, acc);
          umma_bf16_2sm_p(lead, d_bias, one_a, bd + b_kstep, idesc, 1u);
          umma_bf16_2sm_p(lead, d_bias, one_a, bd + 2 * b_kstep, idesc, 1u);
          umma_bf16_2sm_p(lead, d_bias, one_a, bd + 3 * b_kstep, idesc, 1u);
        }
        umma_commit_2sm_p(lead, smem_u32(&bars->empty[stage]), 3);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (nblk > 0) umma_commit_2sm_p(elect_pred(), smem_u32(&bars->acc_full), 3);
      __syncwarp();
      wp_stamp(p, 5, lane);
    }
  } else {
    // ======================= epilogue (both CTAs, each drains its own 128 accumulator rows) =======================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;                                    // the two warps of a lane quarter take alternate taps
    const int m = mp * 256 + rank * 128 + quarter * 32 + lane;   // wide channel of this thread's accumulator row
    const bool m_ok = m < p.Cw;
    if (nblk > 0) {
      mbar_wait(smem_u32(&bars->acc_full), 0);
      tc_fence_after();
    }
    if (warp == 2) wp_stamp(p, 6, lane);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int nvalid = min(p.n_pair, p.Cn - c0);                 // narrow channels of this slice that exist
    if (p.tma_out) {
      // every MMA has retired, so the operand ring is free: each warp stages its 32 rows x 128 B slabs there
      // (128B-swizzled) and sends them off as tensor stores {32 columns, 32 rows, 1}: full-line writes instead of
      // 32 different lines per store instruction; rows past Cw / columns past Cn are clipped by the tensor map
      // up to four staging slabs per warp: a slab is only rewritten once the store issued nbuf slabs earlier has
      // finished READING it (wait_group.read nbuf-1), so the tensor stores run behind the TMEM drain
      const int nbuf = p.stages * (int)p.stage_bytes >= 8 * 4 * 4096 ? 4 : (p.stages * (int)p.stage_bytes >= 8 * 2 * 4096 ? 2 : 1);
      uint8_t* stg0 = ring + ew * nbuf * 4096;
      const int row0 = mp * 256 + rank * 128 + quarter * 32;
      int buf = 0;
      for (int t = half; t < ntap; t += 2) {
        for (int c = 0; c < nvalid; c += 32) {
          uint32_t v[32];
          if (nblk > 0) {
            tmem_ld32(t_lane + (uint32_t)(t * p.n_pair + c), v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          if (lane == 0) {
            if (nbuf == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            else if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
          uint8_t* stg = stg0 + buf * 4096;
#pragma unroll
          for (int qv = 0; qv < 8; ++qv)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((qv ^ (lane & 7)) << 4)) = make_uint4(v[4 * qv], v[4 * qv + 1], v[4 * qv + 2], v[4 * qv + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmP, smem_u32(stg), c0 + c, row0, split * p.taps + tap0 + t);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (++buf == nbuf) buf = 0;
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else {
      const bool vec4 = false;
      for (int t = half; t < ntap; t += 2) {
        float* dst = p.partial + (((long long)split * p.taps + tap0 + t) * p.Cw + m) * p.Cn + c0;
        for (int c = 0; c < nvalid; c += 32) {
          uint32_t v[32];
          if (nblk > 0) {
            tmem_ld32(t_lane + (uint32_t)(t * p.n_pair + c), v);   // (columns past n_pair are never stored)
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          if (m_ok) {
            const int nc = min(32, nvalid - c);
            for (int j = 0; j < nc; ++j) dst[c + j] = __uint_as_float(v[j]);
          }
          (void)vec4;
        }
      }
    }
    if (do_bias && half == 1) {
      const uint32_t t_bias = t_lane + (uint32_t)(ntap * p.n_pair);
      if (p.bias_mode == 1) {                                    // one value per accumulator row (all 16 columns equal)
        uint32_t v[16];
        if (nblk > 0) { tmem_ld16(t_bias, v); tmem_ld_wait(); } else { v[0] = 0u; }
        if (m_ok) p.bias_partial[(long long)split * p.Cw + m] = __uint_as_float(v[0]);
      } else if (rank == 0 && quarter == 0) {                    // every row holds the column sums: row 0 stores them
        for (int c = 0; c < nvalid; c += 32) {
          uint32_t v[32];
          if (nblk > 0) {
            tmem_ld32(t_bias + (uint32_t)c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          if (lane == 0) {
            const int nc = min(32, nvalid - c);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nc) p.bias_partial[(long long)split * p.Cn + c0 + c + j] = __uint_as_float(v[j]);
          }
        }
      }
    }
    if (warp == 2) wp_stamp(p, 7, lane);
    tc_fence_before();
  }
  tc_fence_before();
  cluster_sync_all();            // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
struct WgPairPlan {
  int wide_is_dy, Cw, Cn, n_pair, n_half, n_slices, tap_groups, m_pairs, splits, stages, halo_w;
  int tap_begin[WP_MAX_GROUPS + 1];
  int blocks_h, blocks_w;
  long long num_blocks, blocks_per_split;
  uint32_t stage_bytes, narrow_bytes, tx_bytes;
};

static int wg_halo_w() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SININN_WG_HALO");
    v = (e && atoi(e) == 16) ? 16 : 10;
  }
  return v;
}

bool wgrad_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SININN_WG_PAIR");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

// with_bias: 0 none, else reserve the bias accumulator columns in tap group 0
bool plan_wgrad_pair(const sininn_wgrad_desc* d, bool with_bias, WgPairPlan& w) {
  if (!wgrad_pair_enabled()) return false;
  w.wide_is_dy = d->Cout >= d->Cin ? 1 : 0;
  w.Cw = w.wide_is_dy ? d->Cout : d->Cin;
  w.Cn = w.wide_is_dy ? d->Cin : d->Cout;
  if (w.Cw <= 128) return false;                         // a pair would carry an empty half: single-CTA kernel
  const int n_pad = (w.Cn + 15) / 16 * 16;
  w.n_slices = (n_pad + 127) / 128;
  w.n_pair = ((n_pad + w.n_slices - 1) / w.n_slices + 15) / 16 * 16;
  if (w.n_slices > 1) w.n_pair = (w.n_pair + 31) / 32 * 32;   // slab stores of 32 columns must not cross into the next slice
  w.n_half = w.n_pair / 2;
  if (w.n_half > 64) return false;
  w.m_pairs = (w.Cw + 255) / 256;
  const int bias_cols = !with_bias ? 0 : (w.wide_is_dy ? 16 : w.n_pair);
  const int t_max = TMEM_COLS / w.n_pair;
  int t0 = (TMEM_COLS - bias_cols) / w.n_pair;
  if (t0 < 1) return false;
  if (t0 > d->taps) t0 = d->taps;
  int g = 0, t = 0;
  w.tap_begin[0] = 0;
  while (t < d->taps) {
    const int take = g == 0 ? t0 : (t_max < d->taps - t ? t_max : d->taps - t);
    t += take;
    if (++g > WP_MAX_GROUPS) return false;
    w.tap_begin[g] = t;
  }
  w.tap_groups = g;
  w.halo_w = d->taps == 9 ? wg_halo_w() : WP_BLK;
  w.tx_bytes = WP_WIDE_BYTES + (uint32_t)(d->taps == 9 ? w.halo_w * (WP_BLK + 2) : WP_BLK * WP_BLK) * 128u;
  w.narrow_bytes = ((w.tx_bytes - WP_WIDE_BYTES) + 1023u) & ~1023u;
  w.stage_bytes = WP_WIDE_BYTES + w.narrow_bytes;
  const int budget = 227 * 1024 - (int)WP_ONES_BYTES - (int)sizeof(WgPairBarriers) - 1024;
  w.stages = budget / (int)w.stage_bytes;
  if (w.stages > MAX_STAGES) w.stages = MAX_STAGES;
  if (w.stages < 2) return false;
  w.blocks_h = (d->H + WP_BLK - 1) / WP_BLK;
  w.blocks_w = (d->W + WP_BLK - 1) / WP_BLK;
  w.num_blocks = (long long)d->B * w.blocks_h * w.blocks_w;
  const long long items = (long long)w.m_pairs * w.n_slices * w.tap_groups;
  const long long pairs = sm_count() / 2;
  long long s = pairs / items;                           // one wave of pairs ...
  if (s > w.num_blocks / 8) s = w.num_blocks / 8;        // ... of at least 8 K steps each (partials cost bandwidth)
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  w.blocks_per_split = (w.num_blocks + s - 1) / s;
  w.splits = (int)((w.num_blocks + w.blocks_per_split - 1) / w.blocks_per_split);
  return true;
}

size_t wgrad_pair_workspace_bytes(const sininn_wgrad_desc* d) {
  WgPairPlan w;
  if (!plan_wgrad_pair(d, true, w)) return 0;
  return (size_t)w.splits * d->taps * d->Cout * d->Cin * sizeof(float) + (size_t)w.splits * d->Cout * sizeof(float);
}

// Launches the pair kernel; on success fills *splits / *bias_rows / *partial / *bias_partial for the reduction launch.
int launch_wgrad_pair(const sininn_wgrad_desc* d, cudaStream_t st, int* splits, int* wide_is_dy, float** partial,
                      float** bias_partial, int* bias_rows) {
  WgPairPlan w;
  if (!plan_wgrad_pair(d, d->dbias != nullptr, w)) return SININN_EUNSUPPORTED;
  const size_t need_w = (size_t)w.splits * d->taps * d->Cout * d->Cin * sizeof(float);
  const size_t need = need_w + (d->dbias ? (size_t)w.splits * d->Cout * sizeof(float) : 0);
  if (!d->workspace || d->workspace_bytes < need) {
    set_error("wgrad_tc(pair): workspace too small (%zu < %zu)", d->workspace_bytes, need);
    return SININN_EWORKSPACE;
  }
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("wgrad_tc(pair): cuTensorMapEncodeTiled not available from the driver");
    return SININN_ECUDA;
  }
  const void* wide = w.wide_is_dy ? d->dy : d->x;
  const void* narrow = w.wide_is_dy ? d->x : d->dy;
  const int wide_stride = w.wide_is_dy ? d->dy_stride : d->x_stride;
  const int narrow_stride = w.wide_is_dy ? d->x_stride : d->dy_stride;
  CUtensorMap tmW, tmN;
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? wide : narrow;
    const int C = which == 0 ? w.Cw : w.Cn;
    const int stride = which == 0 ? wide_stride : narrow_stride;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)stride * 2, (cuuint64_t)d->W * stride * 2, (cuuint64_t)d->H * d->W * stride * 2};
    const bool halo = which == 1 && d->taps == 9;
    cuuint32_t box[4] = {64, (cuuint32_t)(halo ? w.halo_w : WP_BLK), (cuuint32_t)(halo ? WP_BLK + 2 : WP_BLK), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(which == 0 ? &tmW : &tmN, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("wgrad_tc(pair): cuTensorMapEncodeTiled failed with %d (C=%d stride=%d)", (int)r, C, stride);
      return SININN_ECUDA;
    }
  }
  const int tma_out = (w.Cn % 4) == 0 ? 1 : 0;
  CUtensorMap tmP = tmW;
  if (tma_out) {
    cuuint64_t dims[3] = {(cuuint64_t)w.Cn, (cuuint64_t)w.Cw, (cuuint64_t)w.splits * d->taps};
    cuuint64_t strides[2] = {(cuuint64_t)w.Cn * 4, (cuuint64_t)w.Cw * w.Cn * 4};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmP, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d->workspace, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("wgrad_tc(pair): tensor map (partials) failed with %d (Cn=%d Cw=%d)", (int)r, w.Cn, w.Cw);
      return SININN_ECUDA;
    }
  }
  WgPairParams p;
  p.tma_out = tma_out;
  p.trace = g_wg_trace;
  p.B = d->B; p.H = d->H; p.W = d->W; p.taps = d->taps;
  p.narrow_is_x = w.wide_is_dy;
  p.Cw = w.Cw; p.Cn = w.Cn;
  p.n_pair = w.n_pair; p.n_half = w.n_half; p.n_slices = w.n_slices;
  p.tap_groups = w.tap_groups; p.m_pairs = w.m_pairs; p.splits = w.splits;
  for (int i = 0; i <= WP_MAX_GROUPS; ++i) p.tap_begin[i] = i <= w.tap_groups ? w.tap_begin[i] : d->taps;
  p.blocks_h = w.blocks_h; p.blocks_w = w.blocks_w; p.num_blocks = w.num_blocks; p.blocks_per_split = w.blocks_per_split;
  p.stages = w.stages; p.halo_w = w.halo_w;
  p.stage_bytes = w.stage_bytes; p.narrow_bytes = w.narrow_bytes; p.tx_bytes = w.tx_bytes;
  p.partial = reinterpret_cast<float*>(d->workspace);
  p.bias_mode = d->dbias ? (w.wide_is_dy ? 1 : 2) : 0;
  p.bias_partial = d->dbias ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->workspace) + need_w) : nullptr;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("wgrad_tc(pair): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return SININN_ECUDA;
    }
    attr_set[dev] = true;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + WP_ONES_BYTES + sizeof(WgPairBarriers) + 1024;
  const unsigned grid = 2u * (unsigned)(w.m_pairs * w.n_slices * w.tap_groups * w.splits);
  launch_k(wgrad_pair_kernel, dim3(grid), dim3(WP_THREADS), smem, st, tmW, tmN, tmP, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("wgrad_tc(pair): launch failed: %s", cudaGetErrorString(e));
    return SININN_ECUDA;
  }
  *splits = w.splits;
  *wide_is_dy = w.wide_is_dy;
  *partial = p.partial;
  *bias_partial = p.bias_partial;
  *bias_rows = d->dbias ? w.splits : 0;
  return SININN_OK;
}

}  // namespace tc
}  // namespace sininn
