// 3x3 implicit-GEMM convolution on CTA PAIRS (tcgen05 cta_group::2 / TMEM / TMA, sm_100a).
//
// conv_tc_halo.cu re-streams the weight slab of every (64-channel slab, tap) from L2 for every 128-pixel tile; ncu
// shows those kernels pinned at the L2 -> SM bandwidth, not at the tensor pipe.  Here two CTAs of a cluster (one TPC)
// work on two pixel tiles at once with ONE M = 256 tcgen05.mma per K step:
//   * each CTA loads the halo'd activation patch of its OWN pixel tile (A rows 0..127 / 128..255), as in
//     conv_tc_halo.cu (here a box {64 ch, 10 w, 18 h}; every tap a shifted UMMA descriptor), and
//   * each CTA holds only HALF of the weight rows (B is split along N across the pair), so the weight bytes per
//     pixel tile halve -- and when the per-CTA half of ALL taps and slabs fits in shared memory (every level-0
//     shape) the weights are loaded ONCE per kernel and stay resident: L2 -> SM traffic is the activations only.
// Channel counts need not be multiples of 64: TMA zero-fills the missing channels of a slab and the issuer skips
// the all-zero K = 16 steps, so a 24- or 48-channel operand costs 2 or 3 MMAs per tap, not 4.
// Roles per CTA: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2-9 epilogue (tc_epilogue.cuh).
// Barriers: TMA of both CTAs signals the LEADER's full barriers (.cta_group::2 loads); tcgen05.commit multicasts the
// "stage free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps arrive remotely on the leader's
// "accumulator drained" barrier.
//
// Replaces cuDNN's 3x3 nn.Conv2d forward / data-gradient inside subnet_conv (/root/reference/archs.py:11-13) and
// DenseBlock (/root/reference/archs.py:77-81,88-95).
#include <stdlib.h>
#include "tc_epilogue.cuh"

namespace sininn {
namespace tc {

constexpr int PT_W = 8, PT_H = 16;                               // pixel tile of one CTA (UMMA rows 128)
// halo box: 10 pixels wide (tile + 1 on each side), not 16 -- the 128-byte swizzle of TMA and UMMA is a function of the
// shared-memory address bits, so any 128-byte row pitch is a legal 8-row-group stride (SBO = 1280 B; verified on the B200 by
// the weight-gradient kernel first).  The level-0 convolutions are bound by what an SM takes in from L2 per tile (~22 B/clk):
// 180 instead of 288 halo pixels per 128 output pixels.
constexpr int PH_W = PT_W + 2, PH_H = PT_H + 2;
constexpr uint32_t PHALO_TX = PH_W * PH_H * 128;                 // 23040 bytes delivered per box
constexpr uint32_t PHALO_BYTES = (PHALO_TX + 1023u) & ~1023u;    // stage pitch (stage bases stay 1024-aligned)
constexpr int P_MAX_A = 6, P_MAX_B = 8;
constexpr int PAIR_THREADS = NUM_THREADS + 32;                   // + warp 10: weight-ring producer

struct __align__(8) PairBarriers {
  uint64_t a_full[P_MAX_A], a_empty[P_MAX_A];
  uint64_t b_full[P_MAX_B], b_empty[P_MAX_B];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t w_full;
  uint32_t tmem_base, pad;
};
constexpr int PAIR_BARRIER_BYTES = 512;
static_assert(sizeof(PairBarriers) <= PAIR_BARRIER_BYTES, "pair barrier block too large");

struct PairParams {
  Params p;                    // p.n_tile = output channels per pair tile; p.k_chunks = 64-channel slabs
  int a_stages, b_stages;      // b_stages == 0: weights resident in shared memory
  int n_half;                  // weight rows per CTA (n_tile / 2)
  uint32_t chunk_bytes;        // shared-memory bytes of one (slab, tap) weight chunk of one CTA (multiple of 1024)
  int tps;                     // taps per weight-ring stage (1 or 3): small chunks share a barrier round trip (an
                               // already-complete mbarrier wait costs ~90 cycles, a small-N tap only ~230 of MMA time)
  uint32_t b_region_bytes;     // resident weights or the weight ring
  int ksteps_last;             // K = 16 steps with real channels in the last slab (1..4)
  long long num_ptiles;        // B * tiles_h * tiles_w
  long long num_items;         // ceil(num_ptiles / 2) * n_tiles
  long long* trace;            // debugging aid (sininn_debug_set_trace): clock64 stamps of CTA 0's roles, or NULL
};

constexpr int TRACE_ROLE_WORDS = 512;
__device__ __forceinline__ void trace_stamp(const PairParams& hp, int role, int& idx, int lane) {
#ifdef SININN_PAIR_TRACE
  if (hp.trace != nullptr && blockIdx.x == 0 && lane == 0 && idx < TRACE_ROLE_WORDS) hp.trace[role * TRACE_ROLE_WORDS + idx] = clock64();
  ++idx;
#endif
}
static long long* g_trace_buf = nullptr;

__device__ __forceinline__ void pair_item(const PairParams& hp, long long it, int rank, int& b, int& h0, int& w0, int& n0) {
  const Params& p = hp.p;
  const int nt = (int)(it % p.n_tiles);
  long long pt = (it / p.n_tiles) * 2 + rank;
  n0 = nt * p.n_tile;
  if (pt >= hp.num_ptiles) { b = p.B; h0 = 0; w0 = 0; return; }   // phantom tile of an odd tail: loads zero-fill, stores clip
  const int tw = (int)(pt % p.tiles_w); pt /= p.tiles_w;
  const int th = (int)(pt % p.tiles_h);
  b = (int)(pt / p.tiles_h);
  h0 = th * PT_H; w0 = tw * PT_W;
}

// One 32-column half of a 128-byte bf16 output slab, for the 16-epilogue-warp variant: the two warps of a pair (same TMEM lane
// quarter, hw = 0 / 1) drain columns c + 32 hw .. + 31 of the warp quarter's 32 rows, write their 64-byte halves of the shared
// 128B-swizzled staging rows, meet on a named barrier, and the pair's leader sends the slab off as ONE tensor store.  Only the
// two shapes that need it (launch_conv_pair): bias + ReLU (+ sign bits), or the stored sign-bit mask alone.
template <int TW>
__device__ __forceinline__ void epilogue_half_slab(const Params& p, const CUtensorMap* tmO, uint8_t* stg, const float* bias_c, uint32_t t_col,
                                                   int b, int h0, int w0, int col0, int quarter, int hw, int lane, int pair_bar) {
  const int row = quarter * 32 + lane;
  const int oh = h0 + row / TW, ow = w0 + row % TW;
  const bool row_ok = (oh < p.H) && (ow < p.W) && (b < p.B);
  const long long pix = ((long long)b * p.H + oh) * p.W + ow;
  const int wi = (col0 >> 5) + hw;                        // sign-bit word of this warp's 32 columns
  uint32_t mb = 0xffffffffu;
  if (p.bits_in != nullptr) mb = (row_ok && wi < p.bit_words) ? __ldg(p.bits_in + pix * p.bit_words + wi) : 0u;
  uint32_t v[32];
  tmem_ld32(t_col + 32 * hw, v);
  tmem_ld_wait();
  uint32_t pk[16];
  uint32_t sign = 0u;
  if (p.act == SININN_ACT_RELU) {
    const float* bsl = bias_c + 32 * hw;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 ba = *reinterpret_cast<const float4*>(bsl + 4 * q);
      pk[2 * q] = bf16x2_relu(pack_bf16(__uint_as_float(v[4 * q]) + ba.x, __uint_as_float(v[4 * q + 1]) + ba.y));
      pk[2 * q + 1] = bf16x2_relu(pack_bf16(__uint_as_float(v[4 * q + 2]) + ba.z, __uint_as_float(v[4 * q + 3]) + ba.w));
    }
    if (p.bits_out != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) sign |= bf16x2_gt0_mask(pk[j]) & ((1u << j) | (1u << (16 + j)));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) & sign_bits_expand(mb, j);
  }
  if (hw == 0 && lane == 0) bulk_wait_read0();             // the pair's previous tensor store has finished reading the staging rows
  asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 * hw + q) ^ (lane & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  if (p.bits_out != nullptr && row_ok && wi < p.bit_words) p.bits_out[pix * p.bit_words + wi] = sign;
  fence_async_smem();                                      // generic-proxy smem writes -> visible to the TMA engine
  asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
  if (hw == 0 && lane == 0) {
    tma_store_4d(tmO, smem_u32(stg), col0, w0, h0 + (32 / TW) * quarter, b);
    bulk_commit();
  }
}

// EW = epilogue warps per CTA: 8 (two per TMEM lane quarter, every epilogue variant) or 16 (four per quarter, the 256-wide bf16
// outputs whose epilogue -- not the MMAs -- sets the tile period: with two warps per scheduler the ALU pipe sat at 16 %).
template <int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW + 32, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const PairParams hp) {
  constexpr int RING_WARP = 2 + EW;                        // weight-ring producer
  const Params& p = hp.p;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* b_reg = smem_raw + pad;                                           // resident weights / weight ring
  uint8_t* a_ring = b_reg + hp.b_region_bytes;                               // [a_stages][PHALO_BYTES]
  uint8_t* staging = a_ring + (size_t)hp.a_stages * PHALO_BYTES;             // 1024-aligned
  PairBarriers* bars = reinterpret_cast<PairBarriers*>(staging + NUM_EPI_WARPS * STAGING_BYTES);
  float* bias_s = reinterpret_cast<float*>(staging + NUM_EPI_WARPS * STAGING_BYTES + PAIR_BARRIER_BYTES);

  // warp index through a shuffle: tells ptxas the role branches below are warp-uniform, which lets it keep the
  // MMA/TMA issue loops on the uniform datapath (without it every tcgen05.mma operand costs an R2UR move)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t a_u32 = smem_u32(a_ring), b_u32 = smem_u32(b_reg);
  const int rank = (int)cluster_ctarank();
  const bool resident = hp.b_stages == 0;
  const long long pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P_MAX_A; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 1);
      mbar_init(smem_u32(&bars->a_empty[s]), 1);
    }
    for (int s = 0; s < P_MAX_B; ++s) {
      mbar_init(smem_u32(&bars->b_full[s]), 1);
      mbar_init(smem_u32(&bars->b_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bars->acc_full[a]), 1);
      mbar_init(smem_u32(&bars->acc_empty[a]), 2 * EW);              // the epilogue warps of BOTH CTAs
    }
    mbar_init(smem_u32(&bars->w_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {   // one warp of EACH CTA of the pair allocates (cta_group::2)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();          // barrier inits of both CTAs visible before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();       // prologue above overlaps the previous kernel's tail; no global memory is touched before this
  pdl_trigger();

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    // full barriers live in the leader; this CTA's loads signal them through their shared::cluster address
    const uint32_t w_full_l = mapa_u32(smem_u32(&bars->w_full), 0);
    int sa = 0; uint32_t pa = 0;
    bool w_loaded = false;
    int ti = 0;
    trace_stamp(hp, 0, ti, lane);
    for (long long it = pair; it < hp.num_items; it += npairs) {
      int b, h0, w0, n0;
      pair_item(hp, it, rank, b, h0, w0, n0);
      if (resident && !w_loaded) {         // the pair keeps one N tile for the whole kernel (host guarantees it)
        if (elect_one()) {
          const int nchunks = p.k_chunks * 9;
          if (rank == 0) mbar_expect_tx(smem_u32(&bars->w_full), 2u * (uint32_t)nchunks * (uint32_t)hp.n_half * 128u);
          for (int c = 0; c < nchunks; ++c)
            tma_load_3d_2sm(b_u32 + c * hp.chunk_bytes, &tmB, w_full_l, (c / 9) * 64, n0 + rank * hp.n_half, c % 9);
        }
        __syncwarp();
        w_loaded = true;
      }
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(smem_u32(&bars->a_empty[sa]), pa ^ 1);
        trace_stamp(hp, 0, ti, lane);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(smem_u32(&bars->a_full[sa]), 2u * PHALO_TX);
          tma_load_4d_2sm(a_u32 + sa * PHALO_BYTES, &tmA, mapa_u32(smem_u32(&bars->a_full[sa]), 0), kc * 64, w0 - 1, h0 - 1, b);
        }
        __syncwarp();
        if (++sa == hp.a_stages) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp == RING_WARP) {
    // ======================= weight-ring producer (both CTAs; streaming mode only) =======================
    // its own warp so that the halo loads (one per slab, ~2500 cycles to land) run a_stages slabs ahead instead of
    // being paced by the much shallower weight ring
    if (!resident) {
      int sb = 0; uint32_t pb = 0;
      for (long long it = pair; it < hp.num_items; it += npairs) {
        const int n0 = (int)(it % p.n_tiles) * p.n_tile;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int tap = 0; tap < 9; tap += hp.tps) {
            mbar_wait(smem_u32(&bars->b_empty[sb]), pb ^ 1);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(smem_u32(&bars->b_full[sb]), 2u * (uint32_t)hp.tps * (uint32_t)hp.n_half * 128u);
              const uint32_t full_l = mapa_u32(smem_u32(&bars->b_full[sb]), 0);
              for (int j = 0; j < hp.tps; ++j)
                tma_load_3d_2sm(b_u32 + (sb * hp.tps + j) * hp.chunk_bytes, &tmB, full_l, kc * 64, n0 + rank * hp.n_half, tap + j);
            }
            __syncwarp();
            if (++sb == hp.b_stages) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ======================= MMA issuer (leader CTA; whole warp loops, one elected lane issues) =======================
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 over the pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pb = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint64_t a_desc0 = make_desc(a_u32, PH_W * 128u, 2);
      const uint64_t b_desc0 = make_desc(b_u32, 1024, 2);
      const uint32_t a_step = PHALO_BYTES >> 4, b_step = hp.chunk_bytes >> 4;
      int ti = 0;
      trace_stamp(hp, 1, ti, lane);
      if (resident && pair < hp.num_items) {
        mbar_wait(smem_u32(&bars->w_full), 0);
        tc_fence_after();
      }
      trace_stamp(hp, 1, ti, lane);
      for (long long it = pair; it < hp.num_items; it += npairs) {
        mbar_wait(smem_u32(&bars->acc_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        trace_stamp(hp, 1, ti, lane);
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(smem_u32(&bars->a_full[sa]), pa);
          tc_fence_after();
          trace_stamp(hp, 1, ti, lane);
          __syncwarp();
          const uint32_t lead = elect_pred();
          const uint64_t a_stage = a_desc0 + (uint64_t)(sa * a_step);
          const uint64_t b_slab = b_desc0 + (uint64_t)(kc * 9 * b_step);
          const int ksteps = (kc == p.k_chunks - 1) ? hp.ksteps_last : 4;
          const bool three = hp.tps == 3;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint64_t bdesc;
            if (resident) {
              bdesc = b_slab + (uint64_t)(tap * b_step);
            } else {
              // (tap is a compile-time constant of the unrolled loop: no runtime division)
              if (!three || tap % 3 == 0) {
                mbar_wait(smem_u32(&bars->b_full[sb]), pb);
                tc_fence_after();
              }
              bdesc = b_desc0 + (uint64_t)((three ? sb * 3 + tap % 3 : sb) * b_step);
            }
            // tap (dy, dx) in 0..2 (halo origin is (h0-1, w0-1)): start row dy*PH_W + dx of the halo box, 8 x 16 B per row
            const uint64_t adesc = a_stage + (uint64_t)(((tap / 3) * PH_W + (tap % 3)) * 8);
            umma_bf16_2sm_p(lead, d_tmem, adesc, bdesc, idesc, (kc | tap) != 0 ? 1u : 0u);
            if (ksteps > 1) umma_bf16_2sm_p(lead, d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
            if (ksteps > 2) umma_bf16_2sm_p(lead, d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
            if (ksteps > 3) umma_bf16_2sm_p(lead, d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
            if (!resident && (!three || tap % 3 == 2)) {
              umma_commit_2sm_p(lead, smem_u32(&bars->b_empty[sb]), 3);
              if (++sb == hp.b_stages) { sb = 0; pb ^= 1; }
            }
          }
          umma_commit_2sm_p(lead, smem_u32(&bars->a_empty[sa]), 3);   // halo buffers of both CTAs free
          if (++sa == hp.a_stages) { sa = 0; pa ^= 1; }
        }
        umma_commit_2sm_p(elect_pred(), smem_u32(&bars->acc_full[acc]), 3);
        trace_stamp(hp, 1, ti, lane);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (EW == 16 && warp < RING_WARP) {
    // ======================= epilogue, 16 warps: pairs of warps share a 128-byte slab (epilogue_half_slab) =======================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int sub = ew >> 2;                                 // 0..3
    const int pr = sub >> 1, hw = sub & 1;                   // pair of this quarter, half of the pair's slab
    uint8_t* stg = staging + (quarter * 2 + pr) * STAGING_BYTES;
    const int pair_bar = 2 + quarter * 2 + pr;               // named barriers 2..9 (0 = __syncthreads, 1 = bias load)
    const int etid = threadIdx.x - 64;
    if (etid < 256) bias_s[etid] = (p.bias != nullptr && etid < p.Cout) ? __ldg(p.bias + etid) : 0.f;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    int acc = 0; uint32_t acc_phase = 0;
    for (long long it = pair; it < hp.num_items; it += npairs) {
      int b, h0, w0, n0;
      pair_item(hp, it, rank, b, h0, w0, n0);
      mbar_wait(smem_u32(&bars->acc_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE;
      const int n_valid = min(p.n_tile, p.Cout - n0);
      const int n_slabs = (n_valid + 63) / 64;
      for (int sl = pr; sl < n_slabs; sl += 2)
        epilogue_half_slab<PT_W>(p, &tmO, stg, bias_s + n0 + sl * 64, t_base + sl * 64, b, h0, w0, n0 + sl * 64, quarter, hw, lane, pair_bar);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->acc_empty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_all();
  } else if (EW == 8 && warp < RING_WARP) {
    // ======================= epilogue (both CTAs, each on its own 128 accumulator rows) =======================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    uint8_t* stg = staging + ew * STAGING_BYTES;
    const int etid = threadIdx.x - 64;
    int acc = 0; uint32_t acc_phase = 0;
    int bias_n0 = -1;
    // Cout <= 256: the whole bias vector is loaded once.  Reloading the slice of every N tile costs two 256-thread barriers per
    // work item, which make the eight epilogue warps start every tile in lockstep with the slowest one (in-kernel stamps:
    // ~1000 of the ~3700 cycles a level-0 24->256 tile took)
    const bool bias_all = p.Cout <= 256;
    if (bias_all) epilogue_load_bias(p, bias_s, 0, etid, bias_n0);
    int ti = 0;
    const int tl = warp == 2 ? lane : 1;
    trace_stamp(hp, 2, ti, tl);
    for (long long it = pair; it < hp.num_items; it += npairs) {
      int b, h0, w0, n0;
      pair_item(hp, it, rank, b, h0, w0, n0);
      if (!bias_all) epilogue_load_bias(p, bias_s, n0, etid, bias_n0);
      const float* bias_t = bias_all ? bias_s + n0 : bias_s;      // bias of this tile's first column
      bool row_ok = false; long long pix = 0;
      CplRegs cpl;
      if (p.cpl.mode != 0) {                 // the coupling operands of this thread's first slab travel while the MMAs run
        cpl_row<PT_W>(p, b, h0, w0, quarter, lane, row_ok, pix);
        cpl_prefetch(p.cpl, cpl, pix, half * 16, row_ok);
      }
      mbar_wait(smem_u32(&bars->acc_full[acc]), acc_phase);
      tc_fence_after();
      trace_stamp(hp, 2, ti, tl);
      const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE;
      if (p.cpl.mode != 0) {
        epilogue_tile_coupling(p.cpl, bias_t, t_base, row_ok, pix, half, 2, cpl);
      } else {
#ifdef SININN_PAIR_TRACE
      // stamps inside the first two tiles of warp 2 of CTA 0 go to role slot 2 from word 256 on
      long long* et = (hp.trace != nullptr && blockIdx.x == 0 && warp == 2 && ti < 6) ? hp.trace + 2 * TRACE_ROLE_WORDS + 256 + 16 * (ti / 2) : nullptr;
      epilogue_tile<PT_W>(p, &tmO, stg, bias_t, t_base, b, h0, w0, n0, quarter, half, lane, et);
#else
      epilogue_tile<PT_W>(p, &tmO, stg, bias_t, t_base, b, h0, w0, n0, quarter, half, lane);
#endif
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->acc_empty[acc]), 0));
      trace_stamp(hp, 2, ti, tl);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_out && lane == 0) bulk_wait_all();
    trace_stamp(hp, 2, ti, tl);
  }

  tc_fence_before();
  cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static bool pair_epi16_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SININN_PAIR_EPI16");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

// Host launcher, called by sininn_conv_tc for 3x3 convolutions.  Returns SININN_EUNSUPPORTED when the shape does
// not qualify (the caller then uses the single-CTA kernels).
int launch_conv_pair(const sininn_conv_desc* d, Params p, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("conv_tc(pair): cuTensorMapEncodeTiled not available from the driver");
    return SININN_ECUDA;
  }
  if (d->taps != 9 || (d->rows_pad % 16) != 0) return SININN_EUNSUPPORTED;
  const int sms = sm_count();
  const int npairs = sms / 2;
  if (npairs < 1) return SININN_EUNSUPPORTED;
  PairParams hp;
  p.kc = 64;
  p.k_chunks = (d->Cin + 63) / 64;
  hp.ksteps_last = ((d->Cin - (p.k_chunks - 1) * 64) + 15) / 16;
  p.tiles_h = (d->H + PT_H - 1) / PT_H;
  p.tiles_w = (d->W + PT_W - 1) / PT_W;
  hp.num_ptiles = (long long)d->B * p.tiles_h * p.tiles_w;
  const int fixed = (int)(NUM_EPI_WARPS * STAGING_BYTES) + PAIR_BARRIER_BYTES + 1024 /*bias*/ + 1024 /*alignment*/;
  const int total = 227 * 1024;
  auto chunk_of = [](int n_tile) { return (uint32_t)(((n_tile / 2) * 128 + 1023) & ~1023); };
  // 1. weights resident: whole N (<= 256), or N split in 128-channel tiles when the pair count divides evenly
  int n_tile = 0, a_stages = 2, b_stages = 0, tps = 1;
  {
    int cands[2] = {d->rows_pad <= 256 ? d->rows_pad : 0, (d->rows_pad > 128 && d->rows_pad % 128 == 0) ? 128 : 0};
    for (int ci = 0; ci < 2 && n_tile == 0; ++ci) {
      const int nt = cands[ci];
      if (nt == 0) continue;
      const int n_tiles = d->rows_pad / nt;
      if (npairs % n_tiles != 0) continue;
      // resident weights must leave room for >= 3 halo stages (one halo load takes ~2500 cycles to land)
      const long long wres = (long long)9 * p.k_chunks * chunk_of(nt);
      const long long room = total - fixed - wres;
      if (room >= 3LL * PHALO_BYTES) {
        n_tile = nt;
        a_stages = (int)(room / PHALO_BYTES);
        if (a_stages > P_MAX_A) a_stages = P_MAX_A;
      }
    }
  }
  if (n_tile == 0) {   // 2. weights streamed through a ring, half of every chunk per CTA
    n_tile = d->rows_pad <= 256 ? d->rows_pad : 256;
    if (d->rows_pad % n_tile != 0) return SININN_EUNSUPPORTED;
    // the weight ring comes first (a tap consumes a chunk every few hundred cycles and a chunk takes ~1500 cycles to
    // land: 8 chunks in flight), the halo stages (one per slab) get what is left: 2..4
    tps = chunk_of(n_tile) <= 8192 ? 3 : 1;
    const int chunk = (int)chunk_of(n_tile) * tps;               // bytes of one ring stage
    const int want = tps == 3 ? 4 : P_MAX_B;
    a_stages = 4;
    while (a_stages > 2 && (total - fixed - a_stages * (int)PHALO_BYTES) / chunk < want) --a_stages;
    b_stages = (total - fixed - a_stages * (int)PHALO_BYTES) / chunk;
    if (b_stages > P_MAX_B) b_stages = P_MAX_B;
    if (b_stages < 3) return SININN_EUNSUPPORTED;
  }
  p.n_tile = n_tile;
  p.n_tiles = d->rows_pad / n_tile;
  hp.n_half = n_tile / 2;
  hp.chunk_bytes = chunk_of(n_tile);
  hp.a_stages = a_stages; hp.b_stages = b_stages; hp.tps = tps;
  hp.b_region_bytes = b_stages == 0 ? (uint32_t)(9 * p.k_chunks) * hp.chunk_bytes : (uint32_t)(b_stages * tps) * hp.chunk_bytes;
  hp.num_items = ((hp.num_ptiles + 1) / 2) * p.n_tiles;
  p.num_tiles = hp.num_items;
  hp.p = p;
  hp.trace = g_trace_buf;
  const int esz = p.out_f32 ? 4 : 2;
  CUtensorMap tmA, tmB, tmO;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_stride * 2, (cuuint64_t)d->W * d->in_stride * 2,
                             (cuuint64_t)d->H * d->W * d->in_stride * 2};
    cuuint32_t box[4] = {64, PH_W, PH_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc(pair): tensor map (activations) failed with %d", (int)r); return SININN_ECUDA; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->k_pad, (cuuint64_t)d->rows_pad, (cuuint64_t)d->taps};
    cuuint64_t strides[2] = {(cuuint64_t)d->k_pad * 2, (cuuint64_t)d->rows_pad * d->k_pad * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)hp.n_half, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->wpack), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc(pair): tensor map (weights) failed with %d", (int)r); return SININN_ECUDA; }
  }
  if (p.tma_out) {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->out_stride * esz, (cuuint64_t)d->W * d->out_stride * esz,
                             (cuuint64_t)d->H * d->W * d->out_stride * esz};
    cuuint32_t box[4] = {(cuuint32_t)(128 / esz), PT_W, 32 / PT_W, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmO, p.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d->out, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc(pair): tensor map (output) failed with %d", (int)r); return SININN_ECUDA; }
  } else {
    tmO = tmA;
  }
  const size_t smem = (size_t)hp.b_region_bytes + (size_t)a_stages * PHALO_BYTES + fixed;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_pair_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_pair_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_tc(pair): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SININN_ECUDA; }
    attr_set[dev] = true;
  }
  long long pairs = hp.num_items < npairs ? hp.num_items : npairs;
  if (b_stages == 0 && p.n_tiles > 1) pairs = npairs;            // resident N tile must be invariant per pair
  // 16 epilogue warps where the epilogue sets the pace: wide bf16 outputs that are bias + ReLU (+ sign bits) or the masked data
  // gradient (the two packed-bf16x2 shapes of tc_epilogue.cuh); everything else keeps the general 8-warp epilogue
  const bool relu_shape = p.act == SININN_ACT_RELU && p.bits_in == nullptr;
  const bool mask_shape = p.act == SININN_ACT_NONE && p.bias == nullptr && p.bits_in != nullptr && p.bits_out == nullptr;
  const bool wide16 = pair_epi16_enabled() && p.tma_out && !p.out_f32 && p.cpl.mode == 0 && p.alpha == 1.0f && !p.accumulate && p.mask == nullptr &&
                      d->Cout >= 128 && d->Cout <= 256 && (d->Cout % 64) == 0 && (relu_shape || mask_shape);
  if (wide16) launch_k(conv_tc_pair_kernel<16>, dim3((unsigned)(2 * pairs)), dim3(64 + 32 * 16 + 32), smem, st, tmA, tmB, tmO, hp);
  else launch_k(conv_tc_pair_kernel<8>, dim3((unsigned)(2 * pairs)), dim3(PAIR_THREADS), smem, st, tmA, tmB, tmO, hp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("conv_tc(pair): launch failed: %s", cudaGetErrorString(e)); return SININN_ECUDA; }
  return SININN_OK;
}

void set_pair_trace(long long* buf) { g_trace_buf = buf; }
void set_wgrad_pair_trace(long long* buf);      // wgrad_pair.cu (its 8 stamps go to words 1536.. of the same buffer)
void set_s1bwd_trace(long long* buf);           // subnet1x1_bwd.cu (2 roles x 16 tiles x 8 stamps at words 1600..1855)

}  // namespace tc
}  // namespace sininn

extern "C" int sininn_debug_set_trace(void* device_buf_3x512_int64) {
  sininn::tc::set_pair_trace(reinterpret_cast<long long*>(device_buf_3x512_int64));
  sininn::tc::set_wgrad_pair_trace(device_buf_3x512_int64 ? reinterpret_cast<long long*>(device_buf_3x512_int64) + 3 * 512 : nullptr);
  sininn::tc::set_s1bwd_trace(device_buf_3x512_int64 ? reinterpret_cast<long long*>(device_buf_3x512_int64) + 1600 : nullptr);
  return SININN_OK;
}
