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
  int nterms;                    // split operands: channel-block pairs the K walk is repeated over (1 = plain operands)
  int woff[6], noff[6];          // channel offsets of pair t in the wide / narrow operand
  int bias_terms;                // bit t: pair t's dy block is summed into the bias gradient
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

// Up to four weight gradients share one launch (the two convolutions of a subnet, or the four of a coupling block):
// the pairs of the grid are divided among the problems in proportion to their work, so every problem is cut in
// FEWER pixel splits than it would be alone -- fewer partial tiles to write (the epilogue of a pair is bound by one
// SM's store bandwidth: up to 256 KB of accumulators, ~8000 cycles) and to reduce, and one prologue per pair amortised
// over 2-4x as many K steps.
constexpr int WG_MAX_PROBLEMS = 4;
struct WgTensorMaps { CUtensorMap m[3 * WG_MAX_PROBLEMS]; };      // per problem: wide, narrow, partial
struct WgGroupParams {
  int nprob;
  int pair_begin[WG_MAX_PROBLEMS + 1];
  WgPairParams prob[WG_MAX_PROBLEMS];
};

struct __align__(8) WgPairBarriers {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  uint64_t acc_full;
  uint32_t tmem_base, pad;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WP_THREADS, 1)
wgrad_pair_kernel(const __grid_constant__ WgTensorMaps T, const __grid_constant__ WgGroupParams G) {
  int prob = 0;
  while (prob + 1 < G.nprob && (int)(blockIdx.x >> 1) >= G.pair_begin[prob + 1]) ++prob;
  const WgPairParams& p = G.prob[prob];
  const CUtensorMap& tmW = T.m[3 * prob], & tmN = T.m[3 * prob + 1], & tmP = T.m[3 * prob + 2];
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* ring = smem_raw + pad;
  uint8_t* ones = ring + (size_t)p.stages * p.stage_bytes;                    // 1024-aligned (stage_bytes % 1024 == 0)
  WgPairBarriers* bars = reinterpret_cast<WgPairBarriers*>(ones + WP_ONES_BYTES);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t ring_u32 = smem_u32(ring);
  const int rank = (int)cluster_ctarank();

  // work item of this pair
  int q = (int)(blockIdx.x >> 1) - G.pair_begin[prob];
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
      for (int term = 0; term < p.nterms; ++term) {          // all channel-block pairs of a pixel block back to back
        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(smem_u32(&bars->full[stage]), 2u * p.tx_bytes);
          const uint32_t full_l = mapa_u32(smem_u32(&bars->full[stage]), 0);     // the leader's barrier
          const uint32_t dst = ring_u32 + stage * p.stage_bytes;
          const int wc = wch + p.woff[term], nc = nch + p.noff[term];
          tma_load_4d_2sm(dst, &tmW, full_l, wc, w0, h0, b);
          tma_load_4d_2sm(dst + WP_WIDE_BYTES / 2, &tmW, full_l, wc + 64, w0, h0, b);
          tma_load_4d_2sm(dst + WP_WIDE_BYTES, &tmN, full_l, nc, w0 - ho, h0 - ho, b);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
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
      int stage = 0; uint32_t phase = 0;
      const long long nsteps = nblk * p.nterms;
      int term = 0;
      uint32_t bias_acc = 0u;
      for (long long i = 0; i < nsteps; ++i) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        tc_fence_after();
        if (i == 0) wp_stamp(p, 4, lane);
        __syncwarp();
        const uint32_t lead = elect_pred();
        const uint64_t ad = a_desc0 + (uint64_t)(stage * stage_step);
        const uint64_t bs = b_desc0 + (uint64_t)(stage * stage_step);
        const uint32_t acc = i != 0 ? 1u : 0u;
        const int bias_mode = (do_bias && ((p.bias_terms >> term) & 1)) ? p.bias_mode : 0;
        if (++term == p.nterms) term = 0;
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
          umma_bf16_2sm_p(lead, d_bias, ad, one_b, idesc_b1, bias_acc);
          umma_bf16_2sm_p(lead, d_bias, ad + 128, one_b, idesc_b1, 1u);
          umma_bf16_2sm_p(lead, d_bias, ad + 256, one_b, idesc_b1, 1u);
          umma_bf16_2sm_p(lead, d_bias, ad + 384, one_b, idesc_b1, 1u);
        } else if (bias_mode == 2) {          // colsum of the narrow operand (unshifted): D[*][n] += sum_k 1 * Narrow[k][n]
          const uint64_t bd = bs + center_off;
          umma_bf16_2sm_p(lead, d_bias, one_a, bd, idesc, bias_acc);
          umma_bf16_2sm_p(lead, d_bias, one_a, bd + b_kstep, idesc, 1u);
          umma_bf16_2sm_p(lead, d_bias, one_a, bd + 2 * b_kstep, idesc, 1u);
          umma_bf16_2sm_p(lead, d_bias, one_a, bd + 3 * b_kstep, idesc, 1u);
        }
        if (bias_mode != 0) bias_acc = 1u;
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

// ---------------------------------------------------------------- fixed-order reduction of the partials
// dw[co][ci][tap] (+)= sum_split partial[split][tap][m][n]; (co,ci) = (m,n) when the wide operand is dy, else (n,m).
// Fixed summation order => bit-reproducible gradients.  RG threads share one output vector: thread g sums splits
// g, g+RG, ... (independent loads in flight), the RG partial sums are combined in a fixed order through shared memory.
// One launch serves all problems of a group (blocks [block_begin, next block_begin) belong to a job).
struct WgReduceJob {
  const float* partial; int splits, taps, Cw, Cn, wide_is_dy, Cout, Cin;
  float* dw; int accumulate;
  const float* bias_partial; int bias_rows; float* dbias; int dbias_accumulate;
  int block_begin, vec, ru;    // ru: output vectors per thread and pass (RU, or 1 for small problems: more blocks)
  // merged problems (sininn_wgrad_desc.nseg): output rows [row0, row0 + rows) of segment s go to its own dw / dbias
  int nseg;
  struct { int row0, rows, cin; float* dw; int accumulate; float* dbias; int dbias_accumulate; } seg[8];
};

// destination of gradient entry (co, ci, tap) of a merged problem: NULL if no segment wants it
__device__ __forceinline__ float* seg_dst(const WgReduceJob& j, int co, int ci, int tap, int& acc) {
  for (int s = 0; s < j.nseg; ++s) {
    const int r = co - j.seg[s].row0;
    if (r >= 0 && r < j.seg[s].rows) {
      if (ci >= j.seg[s].cin) return nullptr;
      acc = j.seg[s].accumulate;
      return j.seg[s].dw + ((long long)r * j.seg[s].cin + ci) * j.taps + tap;
    }
  }
  return nullptr;
}
struct WgReduceParams { int njobs; WgReduceJob job[WG_MAX_PROBLEMS + 1]; };     // job[njobs].block_begin = grid size

constexpr int RG = 8;          // split groups (one warp each)
constexpr int RU_MAX = 4;      // output vectors per thread and pass: RU x ceil(splits / RG) independent loads in flight
template <int VEC, int RU>
__device__ __forceinline__ void reduce_job(const WgReduceJob& j, int bid, int nblocks, float (*red)[32 * RU_MAX][5]) {
  if (j.bias_partial != nullptr) {
    // per-split column sums of dy: one warp per channel (lane l adds rows l, l+32, ... in order, then a fixed xor tree)
    const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31;
    for (int c = bid * 8 + wid; c < j.Cout; c += nblocks * 8) {
      float s = 0.f;
      for (int k = ln; k < j.bias_rows; k += 32) s += __ldcs(j.bias_partial + (long long)k * j.Cout + c);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (ln == 0) {
        if (j.nseg == 0) {
          j.dbias[c] = j.dbias_accumulate ? j.dbias[c] + s : s;
        } else {
          for (int q = 0; q < j.nseg; ++q) {
            const int r = c - j.seg[q].row0;
            if (r >= 0 && r < j.seg[q].rows && j.seg[q].dbias != nullptr)
              j.seg[q].dbias[r] = j.seg[q].dbias_accumulate ? j.seg[q].dbias[r] + s : s;
          }
        }
      }
    }
  }
  const long long per = (long long)j.taps * j.Cw * j.Cn;
  const long long perv = per / VEC;
  const int o = threadIdx.x & 31, g = threadIdx.x >> 5;            // output slot within the block, split group
  for (long long base = (long long)bid * (32 * RU); base < perv; base += (long long)nblocks * (32 * RU)) {
    float s[RU][VEC];
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int e = 0; e < VEC; ++e) s[u][e] = 0.f;
#pragma unroll 4
    for (int k = g; k < j.splits; k += RG) {
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const long long iv = base + u * 32 + o;
        if (iv < perv) {
          if (VEC == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(j.partial + k * per + iv * VEC));
            s[u][0] += t.x; s[u][VEC > 1 ? 1 : 0] += t.y; s[u][VEC > 2 ? 2 : 0] += t.z; s[u][VEC > 3 ? 3 : 0] += t.w;
          } else {
            s[u][0] += j.partial[k * per + iv];
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int e = 0; e < VEC; ++e) red[g][u * 32 + o][e] = s[u][e];
    __syncthreads();
    // the 128 output vectors of the pass are finished by 128 threads (warps 0..3), fixed order over the split groups
    if (threadIdx.x < 32 * RU) {
      const int slot = threadIdx.x;
      const long long iv = base + slot;
      if (iv < perv) {
        const long long idx = iv * VEC;
        const int n0 = (int)(idx % j.Cn);
        long long r = idx / j.Cn;
        const int m = (int)(r % j.Cw);
        const int tap = (int)(r / j.Cw);
        float old[VEC];
        float* dst[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int n = n0 + e;
          const int co = j.wide_is_dy ? m : n, ci = j.wide_is_dy ? n : m;
          int acc = j.accumulate;
          dst[e] = j.nseg == 0 ? j.dw + ((long long)co * j.Cin + ci) * j.taps + tap : seg_dst(j, co, ci, tap, acc);
          old[e] = (acc && dst[e] != nullptr) ? *dst[e] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float t = red[0][slot][e];
#pragma unroll
          for (int qq = 1; qq < RG; ++qq) t += red[qq][slot][e];
          if (dst[e] != nullptr) *dst[e] = old[e] + t;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgReduceParams R) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[RG][32 * RU_MAX][5];
  int ji = 0;
  while (ji + 1 < R.njobs && (int)blockIdx.x >= R.job[ji + 1].block_begin) ++ji;
  const WgReduceJob& j = R.job[ji];
  const int bid = blockIdx.x - j.block_begin, nblocks = R.job[ji + 1].block_begin - j.block_begin;
  if (j.vec == 4) { if (j.ru == 1) reduce_job<4, 1>(j, bid, nblocks, red); else reduce_job<4, RU_MAX>(j, bid, nblocks, red); }
  else reduce_job<1, RU_MAX>(j, bid, nblocks, red);
}

static void reduce_add_job(WgReduceParams& R, const sininn_wgrad_desc* d, const float* partial, int splits, int wide_is_dy,
                           const float* bias_partial, int bias_rows, int max_blocks) {
  WgReduceJob& j = R.job[R.njobs];
  j.partial = partial; j.splits = splits; j.taps = d->taps;
  j.wide_is_dy = wide_is_dy;
  j.Cw = wide_is_dy ? d->Cout : d->Cin; j.Cn = wide_is_dy ? d->Cin : d->Cout;
  j.Cout = d->Cout; j.Cin = d->Cin;
  j.dw = d->dw; j.accumulate = d->accumulate;
  j.bias_partial = bias_partial; j.bias_rows = bias_rows; j.dbias = d->dbias; j.dbias_accumulate = d->dbias_accumulate;
  j.nseg = d->nseg;
  for (int q = 0; q < d->nseg && q < 8; ++q) {
    j.seg[q].row0 = d->seg[q].row0; j.seg[q].rows = d->seg[q].rows; j.seg[q].cin = d->seg[q].cin;
    j.seg[q].dw = d->seg[q].dw; j.seg[q].accumulate = d->seg[q].accumulate;
    j.seg[q].dbias = d->seg[q].dbias; j.seg[q].dbias_accumulate = d->seg[q].dbias_accumulate;
  }
  j.vec = (j.Cn % 4) == 0 ? 4 : 1;
  const long long per = (long long)d->taps * j.Cw * j.Cn;
  // small problems (the 1x1 convolutions): one output vector per thread, so that the grid still covers the device
  j.ru = (j.vec == 4 && per / j.vec < (long long)32 * RU_MAX * sm_count()) ? 1 : RU_MAX;
  const int RU = j.ru;
  long long g = (per / j.vec + 32 * RU - 1) / (32 * RU);
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  const int begin = R.njobs == 0 ? 0 : R.job[R.njobs].block_begin;
  j.block_begin = begin;
  R.job[R.njobs + 1].block_begin = begin + (int)g;
  ++R.njobs;
}

static int launch_reduce_group(const WgReduceParams& R, cudaStream_t st) {
  launch_k(wgrad_reduce_kernel, dim3((unsigned)R.job[R.njobs].block_begin), dim3(256), 0, st, R);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("wgrad_tc(reduce): launch failed: %s", cudaGetErrorString(e));
    return SININN_ECUDA;
  }
  return SININN_OK;
}

// used by the single-CTA kernel of wgrad_tc.cu
int launch_reduce_single(const sininn_wgrad_desc* d, cudaStream_t st, const float* partial, int splits, int wide_is_dy,
                         const float* bias_partial, int bias_rows) {
  WgReduceParams R;
  R.njobs = 0;
  reduce_add_job(R, d, partial, splits, wide_is_dy, bias_partial, bias_rows, sm_count() * 8);
  return launch_reduce_group(R, st);
}

// used by the fused 1x1 subnet backward (subnet1x1_bwd.cu): its two weight gradients (+ bias gradients) leave the kernel
// as one partial per CTA; both are reduced by one launch.  d0 / d1 carry taps, Cout, Cin, dw, accumulate, dbias,
// dbias_accumulate of the two convolutions; partials are [splits][Cw][Cn], bias partials [splits][Cout].
int launch_reduce_two(const sininn_wgrad_desc* d0, const float* partial0, int wide_is_dy0, const float* bias_partial0,
                      const sininn_wgrad_desc* d1, const float* partial1, int wide_is_dy1, const float* bias_partial1,
                      int splits, cudaStream_t st) {
  WgReduceParams R;
  R.njobs = 0;
  reduce_add_job(R, d0, partial0, splits, wide_is_dy0, bias_partial0, bias_partial0 ? splits : 0, sm_count() * 4);
  reduce_add_job(R, d1, partial1, splits, wide_is_dy1, bias_partial1, bias_partial1 ? splits : 0, sm_count() * 4);
  return launch_reduce_group(R, st);
}

// ---------------------------------------------------------------- host side
struct WgPairPlan {
  int wide_is_dy, Cw, Cn, n_pair, n_half, n_slices, tap_groups, m_pairs, splits, stages, halo_w;
  int tap_begin[WP_MAX_GROUPS + 1];
  int blocks_h, blocks_w;
  long long num_blocks, blocks_per_split;
  uint32_t stage_bytes, narrow_bytes, tx_bytes;
  int items;
  double kstep_cycles;           // rough cost of one K step of one pair (for dividing the grid among problems)
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

// does the problem ask for a bias gradient (its own, or any segment's of a merged problem)?
static bool wants_bias(const sininn_wgrad_desc* d) {
  if (d->nseg == 0) return d->dbias != nullptr;
  for (int q = 0; q < d->nseg; ++q)
    if (d->seg[q].dbias != nullptr) return true;
  return false;
}

static void set_splits(WgPairPlan& w, long long pairs_budget) {
  long long s = pairs_budget / w.items;                  // one wave of pairs ...
  if (s > w.num_blocks / 4) s = w.num_blocks / 4;        // ... of at least 4 K steps each (partials cost bandwidth)
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  w.blocks_per_split = (w.num_blocks + s - 1) / s;
  w.splits = (int)((w.num_blocks + w.blocks_per_split - 1) / w.blocks_per_split);
}

// Tiling of one problem (everything but the number of pixel splits, which depends on the pairs it is given).
static bool plan_wgrad_pair(const sininn_wgrad_desc* d, bool with_bias, WgPairPlan& w) {
  if (!wgrad_pair_enabled()) return false;
  if (d->x_dtype != SININN_BF16 || d->dy_dtype != SININN_BF16 || (d->taps != 1 && d->taps != 9)) return false;
  if (d->nterms < 0 || d->nterms > 6) return false;
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
  w.items = w.m_pairs * w.n_slices * w.tap_groups;
  // measured (tools/wgrad_trace.py): an M=256 MMA costs ~50 cycles for N <= 96 (shared-memory operand reads) and a
  // pair's SMs take in ~22 B/clk each from L2
  const double mma = 4.0 * ((double)d->taps / w.tap_groups + (with_bias ? 1.0 / w.tap_groups : 0.0)) * (w.n_pair > 96 ? w.n_pair / 2.0 : 50.0);
  const double ingest = (double)w.tx_bytes / 22.0;
  w.kstep_cycles = mma > ingest ? mma : ingest;
  return true;
}

bool wgrad_pair_takes(const sininn_wgrad_desc* d) {
  WgPairPlan w;
  return plan_wgrad_pair(d, true, w);
}

size_t wgrad_pair_workspace_bytes(const sininn_wgrad_desc* d) {
  WgPairPlan w;
  if (!plan_wgrad_pair(d, true, w)) return 0;
  set_splits(w, sm_count() / 2);
  return (size_t)w.splits * d->taps * d->Cout * d->Cin * sizeof(float) + (size_t)w.splits * d->Cout * sizeof(float) + 512;
}

// One pair-kernel launch + one reduction launch for n <= WG_MAX_PROBLEMS problems that all qualify for the pair kernel
// (SININN_EUNSUPPORTED otherwise: the caller runs them one by one).  The workspace is carved up here.
int launch_wgrad_pair_group(const sininn_wgrad_desc* ds, int n, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (n < 1 || n > WG_MAX_PROBLEMS) return SININN_EUNSUPPORTED;
  WgPairPlan w[WG_MAX_PROBLEMS];
  const int pairs = sm_count() / 2;
  double work[WG_MAX_PROBLEMS], total = 0.0;
  int items = 0;
  for (int i = 0; i < n; ++i) {
    if (!plan_wgrad_pair(&ds[i], wants_bias(&ds[i]), w[i])) return SININN_EUNSUPPORTED;
    work[i] = (double)w[i].num_blocks * w[i].items * w[i].kstep_cycles * (ds[i].nterms > 0 ? ds[i].nterms : 1);
    total += work[i];
    items += w[i].items;
  }
  if (items > pairs) return SININN_EUNSUPPORTED;
  // pairs in proportion to the work, at least one per item; leftovers go to the problem with the most work per pair
  int given[WG_MAX_PROBLEMS], used = 0;
  for (int i = 0; i < n; ++i) {
    int g = (int)(pairs * work[i] / total);
    g = g / w[i].items * w[i].items;
    if (g < w[i].items) g = w[i].items;
    given[i] = g;
    used += g;
  }
  while (used > pairs) {           // (rounding up the small ones can overshoot)
    int worst = -1;
    for (int i = 0; i < n; ++i)
      if (given[i] > w[i].items && (worst < 0 || work[i] / given[i] < work[worst] / given[worst])) worst = i;
    if (worst < 0) return SININN_EUNSUPPORTED;
    given[worst] -= w[worst].items;
    used -= w[worst].items;
  }
  for (;;) {
    int best = -1;
    for (int i = 0; i < n; ++i)
      if (used + w[i].items <= pairs && (best < 0 || work[i] / given[i] > work[best] / given[best])) best = i;
    if (best < 0) break;
    given[best] += w[best].items;
    used += w[best].items;
  }
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("wgrad_tc(pair): cuTensorMapEncodeTiled not available from the driver");
    return SININN_ECUDA;
  }
  WgTensorMaps T;
  WgGroupParams G;
  WgReduceParams R;
  R.njobs = 0;
  G.nprob = n;
  G.pair_begin[0] = 0;
  size_t off = 0;
  size_t max_smem = 0;
  for (int i = 0; i < n; ++i) {
    const sininn_wgrad_desc* d = &ds[i];
    WgPairPlan& wi = w[i];
    set_splits(wi, given[i]);
    const size_t need_w = (size_t)wi.splits * d->taps * d->Cout * d->Cin * sizeof(float);
    const size_t need_b = wants_bias(d) ? (size_t)wi.splits * d->Cout * sizeof(float) : 0;
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace) + off;
    off += (need_w + need_b + 255) & ~(size_t)255;
    if (!workspace || off > workspace_bytes) {
      set_error("wgrad_tc(pair): workspace too small (%zu < %zu)", workspace_bytes, off);
      return SININN_EWORKSPACE;
    }
    const void* wide = wi.wide_is_dy ? d->dy : d->x;
    const void* narrow = wi.wide_is_dy ? d->x : d->dy;
    const int wide_stride = wi.wide_is_dy ? d->dy_stride : d->x_stride;
    const int narrow_stride = wi.wide_is_dy ? d->x_stride : d->dy_stride;
    for (int which = 0; which < 2; ++which) {
      const void* ptr = which == 0 ? wide : narrow;
      const int stride = which == 0 ? wide_stride : narrow_stride;
      // split operands: the channel extent is the whole row of blocks (the pixel stride); plain: the true channel count
      const int C = d->nterms > 0 ? stride : (which == 0 ? wi.Cw : wi.Cn);
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
      cuuint64_t strides[3] = {(cuuint64_t)stride * 2, (cuuint64_t)d->W * stride * 2, (cuuint64_t)d->H * d->W * stride * 2};
      const bool halo = which == 1 && d->taps == 9;
      cuuint32_t box[4] = {64, (cuuint32_t)(halo ? wi.halo_w : WP_BLK), (cuuint32_t)(halo ? WP_BLK + 2 : WP_BLK), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = encode(&T.m[3 * i + which], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("wgrad_tc(pair): cuTensorMapEncodeTiled failed with %d (C=%d stride=%d)", (int)r, C, stride);
        return SININN_ECUDA;
      }
    }
    const int tma_out = (wi.Cn % 4) == 0 ? 1 : 0;
    T.m[3 * i + 2] = T.m[3 * i];
    if (tma_out) {
      cuuint64_t dims[3] = {(cuuint64_t)wi.Cn, (cuuint64_t)wi.Cw, (cuuint64_t)wi.splits * d->taps};
      cuuint64_t strides[2] = {(cuuint64_t)wi.Cn * 4, (cuuint64_t)wi.Cw * wi.Cn * 4};
      cuuint32_t box[3] = {32, 32, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&T.m[3 * i + 2], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("wgrad_tc(pair): tensor map (partials) failed with %d (Cn=%d Cw=%d)", (int)r, wi.Cn, wi.Cw);
        return SININN_ECUDA;
      }
    }
    WgPairParams& p = G.prob[i];
    p.tma_out = tma_out;
    p.nterms = d->nterms > 0 ? d->nterms : 1;
    p.bias_terms = d->nterms > 0 ? d->bias_term_mask : 1;
    for (int t = 0; t < 6; ++t) {
      const int xo = (d->nterms > 0 && t < d->nterms) ? d->x_term_off[t] : 0, yo = (d->nterms > 0 && t < d->nterms) ? d->dy_term_off[t] : 0;
      p.woff[t] = wi.wide_is_dy ? yo : xo;
      p.noff[t] = wi.wide_is_dy ? xo : yo;
    }
    p.trace = i == 0 ? g_wg_trace : nullptr;
    p.B = d->B; p.H = d->H; p.W = d->W; p.taps = d->taps;
    p.narrow_is_x = wi.wide_is_dy;
    p.Cw = wi.Cw; p.Cn = wi.Cn;
    p.n_pair = wi.n_pair; p.n_half = wi.n_half; p.n_slices = wi.n_slices;
    p.tap_groups = wi.tap_groups; p.m_pairs = wi.m_pairs; p.splits = wi.splits;
    for (int k = 0; k <= WP_MAX_GROUPS; ++k) p.tap_begin[k] = k <= wi.tap_groups ? wi.tap_begin[k] : d->taps;
    p.blocks_h = wi.blocks_h; p.blocks_w = wi.blocks_w; p.num_blocks = wi.num_blocks; p.blocks_per_split = wi.blocks_per_split;
    p.stages = wi.stages; p.halo_w = wi.halo_w;
    p.stage_bytes = wi.stage_bytes; p.narrow_bytes = wi.narrow_bytes; p.tx_bytes = wi.tx_bytes;
    p.partial = reinterpret_cast<float*>(base);
    p.bias_mode = wants_bias(d) ? (wi.wide_is_dy ? 1 : 2) : 0;
    p.bias_partial = wants_bias(d) ? reinterpret_cast<float*>(base + need_w) : nullptr;
    G.pair_begin[i + 1] = G.pair_begin[i] + wi.items * wi.splits;
    const size_t smem = (size_t)p.stages * p.stage_bytes + WP_ONES_BYTES + sizeof(WgPairBarriers) + 1024;
    if (smem > max_smem) max_smem = smem;
    reduce_add_job(R, d, p.partial, wi.splits, wi.wide_is_dy, p.bias_partial, wants_bias(d) ? wi.splits : 0, sm_count() * 8 / n);
  }
  for (int i = n; i < WG_MAX_PROBLEMS; ++i) { G.prob[i] = G.prob[0]; G.pair_begin[i + 1] = G.pair_begin[n]; }
  for (int i = 3 * n; i < 3 * WG_MAX_PROBLEMS; ++i) T.m[i] = T.m[0];
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
  const unsigned grid = 2u * (unsigned)G.pair_begin[n];
  launch_k(wgrad_pair_kernel, dim3(grid), dim3(WP_THREADS), max_smem, st, T, G);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("wgrad_tc(pair): launch failed: %s", cudaGetErrorString(e));
    return SININN_ECUDA;
  }
  return launch_reduce_group(R, st);
}

size_t wgrad_pair_group_workspace_bytes(const sininn_wgrad_desc* ds, int n) {
  // every problem may end up with all the pairs of the device (upper bound of its splits)
  size_t tot = 0;
  for (int i = 0; i < n; ++i) tot += wgrad_pair_workspace_bytes(&ds[i]);     // (0 for problems the pair kernel does not take)
  return tot;
}

}  // namespace tc
}  // namespace sininn
