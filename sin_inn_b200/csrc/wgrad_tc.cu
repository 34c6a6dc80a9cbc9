// tcgen05 weight-gradient kernel (bf16 operands read MN-major straight from the channels-last activations).
#include "tc_common.cuh"

namespace sininn {
int wgrad_simt_splits(const sininn_wgrad_desc* d);
namespace tc {

// wgrad_pair.cu: the CTA-pair kernel (taken whenever the wide operand has more than 128 channels)
size_t wgrad_pair_workspace_bytes(const sininn_wgrad_desc* d);
size_t wgrad_pair_group_workspace_bytes(const sininn_wgrad_desc* ds, int n);
int launch_wgrad_pair_group(const sininn_wgrad_desc* ds, int n, void* workspace, size_t workspace_bytes, cudaStream_t st);
bool wgrad_pair_takes(const sininn_wgrad_desc* d);
int launch_reduce_single(const sininn_wgrad_desc* d, cudaStream_t st, const float* partial, int splits, int wide_is_dy,
                         const float* bias_partial, int bias_rows);

constexpr int NUM_THREADS = 192;
constexpr int SMEM_RING_BUDGET = 200 * 1024;

// ================================================================ weight gradient
//   dW[tap][m][n] = sum_pixels  Wide[p][m] * Narrow[p + off(tap)][n]
// "Wide" is whichever of (x, dy) has more channels (the 256-wide hidden side of every subnet conv); it is the
// M operand in 128-channel tiles and is never shifted.  "Narrow" is the N operand, shifted per tap (by +off when
// it is x, by -off when it is dy).  Both operands are read straight from the channels-last activations as
// MN-major SWIZZLE_128B tiles (K = pixels, 64 channels x 128 B rows): no transposed copies exist anywhere.
// One K step = an 8x8 pixel block.  For 3x3 filters the narrow operand is loaded ONCE per step with its halo
// (box 16 w x 10 h, origin (w0-1, h0-1)) and every tap is a descriptor into that buffer:
//     start = base + ((1+sy)*16 + (1+sx)) * 128 B, K-group (8 pixels of one image row) stride SBO = 2048 B.
// One CTA = (M tile, group of taps, pixel split); the accumulators of all its taps live in TMEM side by side.
// fp32 partials go to the workspace, a fixed-order reduction writes OIHW (deterministic, no float atomics).
constexpr int WG_BLK = 8;                                        // pixel block edge: 64 pixels per K step
constexpr uint32_t WG_BOX_BYTES = WG_BLK * WG_BLK * 128;          // plain 8x8 box: 64 px x 64 ch x 2 B
constexpr int WG_HALO_W = 16, WG_HALO_H = WG_BLK + 2;
constexpr uint32_t WG_HALO_BYTES = WG_HALO_W * WG_HALO_H * 128;   // 20480

struct WgradParams {
  int B, H, W;
  int taps, narrow_is_x;         // shift sign: +off when the narrow operand is x, -off when it is dy
  int Cw, Cn;                    // true channel counts of the wide / narrow operands
  int n_pad, n_groups;           // narrow channels padded to 16; 64-channel boxes per tap
  int taps_per_cta, tap_groups, m_tiles, splits;
  int blocks_h, blocks_w;        // pixel blocks per image
  long long num_blocks, blocks_per_split;
  int stages;
  uint32_t stage_bytes, tx_bytes;
  uint32_t narrow_bytes;         // bytes of one 64-channel narrow box (halo'd for 3x3)
  float* partial;                // [split][tap][Cw][Cn]
  // bias gradient = column sums of dy, computed from the operand tiles the MMA loop streams through shared memory
  // anyway (the four epilogue warps are idle during that loop): 0 off, 1 dy is the wide operand, 2 dy is the narrow one
  int bias_mode;
  float* bias_partial;           // mode 1: [2 * split + row half][Cw]; mode 2: [split][Cn]
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmN, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* ring = smem_raw + pad;
  Barriers* bars = reinterpret_cast<Barriers*>(ring + (size_t)p.stages * p.stage_bytes);
  // warp index through a shuffle: tells ptxas the role branches below are warp-uniform, which lets it keep the
  // MMA/TMA issue loops on the uniform datapath (without it every tcgen05.mma operand costs an R2UR move)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t ring_u32 = smem_u32(ring);

  // work item
  int item = blockIdx.x;
  const int split = item % p.splits; item /= p.splits;
  const int tg = item % p.tap_groups;
  const int mt = item / p.tap_groups;
  const int tap0 = tg * p.taps_per_cta;
  const int ntap = min(p.taps_per_cta, p.taps - tap0);
  const long long blk0 = (long long)split * p.blocks_per_split;
  long long blk1 = blk0 + p.blocks_per_split;
  if (blk1 > p.num_blocks) blk1 = p.num_blocks;
  const long long nblk = blk1 > blk0 ? blk1 - blk0 : 0;

  const bool do_bias = p.bias_mode != 0 && tg == 0 && (p.bias_mode == 1 || mt == 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), do_bias ? 5 : 1);     // MMA commit (+ the four column-sum warps)
    }
    mbar_init(smem_u32(&bars->acc_full[0]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmN);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_wait();       // prologue above overlaps the previous kernel's tail; no global memory is touched before this
  pdl_trigger();

  if (warp == 0) {
    {   // whole warp loops (uniform control flow), one elected lane issues the TMA loads
      int stage = 0; uint32_t phase = 0;
      for (long long i = 0; i < nblk; ++i) {
        long long blk = blk0 + i;
        const int bw = (int)(blk % p.blocks_w); blk /= p.blocks_w;
        const int bh = (int)(blk % p.blocks_h);
        const int b = (int)(blk / p.blocks_h);
        const int w0 = bw * WG_BLK, h0 = bh * WG_BLK;
        mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
        if (elect_one()) {
          const uint32_t full = smem_u32(&bars->full[stage]);
          mbar_expect_tx(full, p.tx_bytes);
          uint32_t dst = ring_u32 + stage * p.stage_bytes;
          tma_load_4d(dst, &tmW, full, mt * 128, w0, h0, b);
          tma_load_4d(dst + WG_BOX_BYTES, &tmW, full, mt * 128 + 64, w0, h0, b);
          dst += 2 * WG_BOX_BYTES;
          const int ho = (p.taps == 9) ? 1 : 0;                  // halo origin offset
          for (int g = 0; g < p.n_groups; ++g) {
            tma_load_4d(dst, &tmN, full, g * 64, w0 - ho, h0 - ho, b);
            dst += p.narrow_bytes;
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {   // whole warp loops, one elected lane issues the MMAs
      // D=f32, A=B=bf16, both MN-major (bits 15, 16), N = n_pad, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      // descriptors built once (MN-major SW128: LBO = distance between 64-channel boxes, SBO = distance between
      // 8-pixel K groups); per MMA only the start-address field (units of 16 B) is advanced
      const bool halo = p.taps == 9;
      uint64_t a_desc0 = make_desc(ring_u32, 1024, 2);
      a_desc0 = (a_desc0 & ~((uint64_t)0x3FFF << 16)) | ((uint64_t)(WG_BOX_BYTES >> 4) << 16);
      uint64_t b_desc0 = make_desc(ring_u32 + 2 * WG_BOX_BYTES, halo ? WG_HALO_W * 128u : 1024u, 2);
      b_desc0 = (b_desc0 & ~((uint64_t)0x3FFF << 16)) | ((uint64_t)(p.narrow_bytes >> 4) << 16);
      const uint32_t b_kstep = (halo ? 2 * WG_HALO_W * 128u : 2048u) >> 4;      // 16 pixels further along K
      uint32_t tap_off[9];                                                       // start row of each tap, in 16 B units
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        int sy = 0, sx = 0;
        if (halo) {
          const int tap = tap0 + t;
          sy = tap / 3 - 1; sx = tap % 3 - 1;
          if (!p.narrow_is_x) { sy = -sy; sx = -sx; }
          sy += 1; sx += 1;
        }
        tap_off[t] = (uint32_t)(sy * WG_HALO_W + sx) * 8u;
      }
      const uint32_t stage_step = p.stage_bytes >> 4;
      for (long long i = 0; i < nblk; ++i) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        tc_fence_after();
        const uint64_t ad = a_desc0 + (uint64_t)(stage * stage_step);
        const uint64_t bs = b_desc0 + (uint64_t)(stage * stage_step);
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            if (t < ntap) {
              const uint64_t bd = bs + tap_off[t];
              const uint32_t d_tmem = tmem_base + t * p.n_pad;
              umma_bf16(d_tmem, ad, bd, idesc, i != 0 ? 1u : 0u);
              umma_bf16(d_tmem, ad + 128, bd + b_kstep, idesc, 1u);
              umma_bf16(d_tmem, ad + 256, bd + 2 * b_kstep, idesc, 1u);
              umma_bf16(d_tmem, ad + 384, bd + 3 * b_kstep, idesc, 1u);
            }
          }
          umma_commit(smem_u32(&bars->empty[stage]));
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(smem_u32(&bars->acc_full[0]));
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int m = mt * 128 + quarter * 32 + lane;            // wide-operand channel of this thread's row
    const bool m_ok = m < p.Cw;
    if (do_bias) {
      // Column sums of dy over this CTA's pixel blocks, read from the MN-major SWIZZLE_128B tiles in the ring
      // (pixel row r = 128 B of 64 channels, 16-byte chunk c stored at c ^ (r & 7)).  Each thread owns a channel
      // pair; within a warp the 32 lanes read 32 different 4-byte words of one row: conflict-free.
      const int et = threadIdx.x - 64;                       // 0..127
      float s0 = 0.f, s1 = 0.f;
      int box, q, r_lo, r_hi;
      uint32_t tile_off;
      if (p.bias_mode == 1) {                                // wide tile: 2 boxes x 64 channels, plain 8x8 pixel boxes
        const int pair = et & 63;
        box = pair >> 5; q = pair & 31;
        r_lo = (et >> 6) * 32; r_hi = r_lo + 32;             // the two thread halves split the 64 pixels
        tile_off = (uint32_t)box * WG_BOX_BYTES;
      } else {                                               // narrow tile: n_groups boxes (halo'd for 3x3)
        box = et >> 5; q = et & 31;
        r_lo = 0; r_hi = 64;
        tile_off = 2 * WG_BOX_BYTES + (uint32_t)box * p.narrow_bytes;
      }
      const bool active = p.bias_mode == 1 || box < p.n_groups;
      const bool halo_rows = p.bias_mode == 2 && p.taps == 9;
      const uint32_t chunk = (uint32_t)(q >> 2), word = (uint32_t)(q & 3) * 4u;
      int stage = 0; uint32_t phase = 0;
      for (long long i = 0; i < nblk; ++i) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        if (active) {
          const uint8_t* tile = ring + (size_t)stage * p.stage_bytes + tile_off;
#pragma unroll 8
          for (int r = r_lo; r < r_hi; ++r) {
            const int row = halo_rows ? ((r >> 3) + 1) * WG_HALO_W + (r & 7) + 1 : r;      // interior pixel of the halo box
            const uint32_t v = *reinterpret_cast<const uint32_t*>(tile + row * 128 + ((chunk ^ (uint32_t)(row & 7)) << 4) + word);
            s0 += __uint_as_float(v << 16);
            s1 += __uint_as_float(v & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->empty[stage]));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (active) {
        if (p.bias_mode == 1) {
          const int ch = mt * 128 + (et & 63) * 2;
          float* dst = p.bias_partial + ((long long)(2 * split + (et >> 6))) * p.Cw + ch;
          if (ch < p.Cw) dst[0] = s0;
          if (ch + 1 < p.Cw) dst[1] = s1;
        } else {
          const int ch = et * 2;
          float* dst = p.bias_partial + (long long)split * p.Cn + ch;
          if (ch < p.Cn) dst[0] = s0;
          if (ch + 1 < p.Cn) dst[1] = s1;
        }
      }
    }
    if (nblk > 0) {
      mbar_wait(smem_u32(&bars->acc_full[0]), 0);
      tc_fence_after();
    }
    const bool vec4 = (p.Cn % 4) == 0;                        // rows of the partial buffer are then 16-byte aligned
    for (int t = 0; t < ntap; ++t) {
      float* dst = p.partial + (((long long)split * p.taps + tap0 + t) * p.Cw + m) * p.Cn;
      for (int c = 0; c < p.Cn; c += 32) {
        uint32_t v[32];
        if (nblk > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + t * p.n_pad + c, v);   // columns past n_pad are never stored
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (m_ok) {
          const int nc = min(32, p.Cn - c);
          if (vec4) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (4 * q < nc)
                *reinterpret_cast<float4*>(dst + c + 4 * q) = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                          __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          } else {
            for (int j = 0; j < nc; ++j) dst[c + j] = __uint_as_float(v[j]);
          }
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

struct WgradPlan {
  int wide_is_dy, Cw, Cn, n_pad, n_groups, taps_per_cta, tap_groups, m_tiles, splits, stages;
  int blocks_h, blocks_w;
  long long num_blocks, blocks_per_split;
  uint32_t stage_bytes, narrow_bytes;
};

static bool plan_wgrad(const sininn_wgrad_desc* d, WgradPlan& w) {
  w.wide_is_dy = d->Cout >= d->Cin ? 1 : 0;
  w.Cw = w.wide_is_dy ? d->Cout : d->Cin;
  w.Cn = w.wide_is_dy ? d->Cin : d->Cout;
  w.n_pad = (w.Cn + 15) / 16 * 16;
  if (w.n_pad > 256) return false;
  w.n_groups = (w.n_pad + 63) / 64;
  w.narrow_bytes = d->taps == 9 ? WG_HALO_BYTES : WG_BOX_BYTES;
  int t = TMEM_COLS / w.n_pad;                       // accumulators of all taps of a CTA sit side by side in TMEM
  if (t > d->taps) t = d->taps;
  w.taps_per_cta = t;
  w.tap_groups = (d->taps + t - 1) / t;
  w.m_tiles = (w.Cw + 127) / 128;
  w.stage_bytes = 2 * WG_BOX_BYTES + w.n_groups * w.narrow_bytes;
  w.stages = SMEM_RING_BUDGET / (int)w.stage_bytes;
  if (w.stages > MAX_STAGES) w.stages = MAX_STAGES;
  if (w.stages < 2) return false;
  w.blocks_h = (d->H + WG_BLK - 1) / WG_BLK;
  w.blocks_w = (d->W + WG_BLK - 1) / WG_BLK;
  w.num_blocks = (long long)d->B * w.blocks_h * w.blocks_w;
  long long items = (long long)w.m_tiles * w.tap_groups;
  long long s = (sm_count() + items - 1) / items;      // about one CTA per SM ...
  if (s > w.num_blocks / 16) s = w.num_blocks / 16;    // ... but at least 16 K steps each (partials cost bandwidth)
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  w.blocks_per_split = (w.num_blocks + s - 1) / s;
  w.splits = (int)((w.num_blocks + w.blocks_per_split - 1) / w.blocks_per_split);
  return true;
}

}  // namespace tc
}  // namespace sininn

using namespace sininn;

extern "C" {

size_t sininn_wgrad_workspace_bytes(const sininn_wgrad_desc* d, int tensor_core) {
  if (!d || d->Cin <= 0 || d->Cout <= 0 || d->taps <= 0) return 0;
  size_t simt = (size_t)wgrad_simt_splits(d) * d->taps * d->Cout * d->Cin * sizeof(float);
  if (!tensor_core) return simt;
  const size_t pairb = sininn::tc::wgrad_pair_workspace_bytes(d);
  if (pairb > simt) simt = pairb;
  sininn::tc::WgradPlan w;
  if (!sininn::tc::plan_wgrad(d, w)) return simt;
  size_t tcb = (size_t)w.splits * d->taps * d->Cout * d->Cin * sizeof(float) + (size_t)2 * w.splits * d->Cout * sizeof(float);
  return tcb > simt ? tcb : simt;
}

size_t sininn_wgrad_group_workspace_bytes(const sininn_wgrad_desc* descs, int n) {
  if (!descs || n < 1) return 0;
  const size_t g = sininn::tc::wgrad_pair_group_workspace_bytes(descs, n);
  size_t single = 0;
  for (int i = 0; i < n; ++i) {
    const size_t b = sininn_wgrad_workspace_bytes(&descs[i], 1);
    if (b > single) single = b;
  }
  return g > single ? g : single;
}

int sininn_wgrad_pair_supported(const sininn_wgrad_desc* d) {
  return (d != nullptr && sininn::tc::wgrad_pair_takes(d)) ? 1 : 0;
}

int sininn_wgrad_tc_group(const sininn_wgrad_desc* descs, int n, void* workspace, size_t workspace_bytes, sininn_stream_t stream) {
  SININN_CHECK_ARG(descs && n >= 1, "wgrad_tc_group: no problems");
  for (int i = 0; i < n; ++i) {
    const sininn_wgrad_desc* d = &descs[i];
    SININN_CHECK_ARG(d->x && d->dy && (d->dw || d->nseg > 0), "wgrad_tc_group: null pointer in problem %d", i);
    SININN_CHECK_ARG(d->nseg >= 0 && d->nseg <= 8, "wgrad_tc_group: at most 8 segments");
    SININN_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "wgrad_tc_group: bad shape in problem %d", i);
    SININN_CHECK_ARG(d->taps == 1 || d->taps == 9, "wgrad_tc_group: taps must be 1 or 9");
    SININN_CHECK_ARG(d->x_dtype == SININN_BF16 && d->dy_dtype == SININN_BF16, "wgrad_tc_group: operands must be bf16");
    SININN_CHECK_ARG(aligned16(d->x) && aligned16(d->dy) && (d->x_stride % 8) == 0 && (d->dy_stride % 8) == 0,
                     "wgrad_tc_group: TMA needs 16-byte aligned operands with pixel strides that are multiples of 8 channels "
                     "(x stride %d, dy stride %d)", d->x_stride, d->dy_stride);
  }
  if (n <= 4) {
    const int rc = sininn::tc::launch_wgrad_pair_group(descs, n, workspace, workspace_bytes, sininn::as_stream(stream));
    if (rc != SININN_EUNSUPPORTED) return rc;
  }
  // Mixed group (e.g. a DenseBlock, archs.py:77-81: its first convolutions have <= 128 input channels): the problems the
  // CTA-pair kernel takes still share one launch, the others run one by one.  Stream order makes sharing the workspace safe.
  sininn_wgrad_desc pairable[4];
  int np = 0;
  for (int i = 0; i < n; ++i) {
    if (np < 4 && sininn::tc::wgrad_pair_workspace_bytes(&descs[i]) > 0) {
      pairable[np++] = descs[i];
      continue;
    }
    sininn_wgrad_desc d = descs[i];
    d.workspace = workspace;
    d.workspace_bytes = workspace_bytes;
    const int rc = sininn_wgrad_tc(&d, stream);
    if (rc != SININN_OK) return rc;
  }
  if (np > 0) {
    const int rc = sininn::tc::launch_wgrad_pair_group(pairable, np, workspace, workspace_bytes, sininn::as_stream(stream));
    if (rc == SININN_EUNSUPPORTED) {
      for (int i = 0; i < np; ++i) {
        pairable[i].workspace = workspace;
        pairable[i].workspace_bytes = workspace_bytes;
        const int r2 = sininn_wgrad_tc(&pairable[i], stream);
        if (r2 != SININN_OK) return r2;
      }
    } else if (rc != SININN_OK) {
      return rc;
    }
  }
  return SININN_OK;
}

int sininn_wgrad_tc(const sininn_wgrad_desc* d, sininn_stream_t stream) {
  using namespace sininn::tc;
  SININN_CHECK_ARG(d && d->x && d->dy && (d->dw || d->nseg > 0), "wgrad_tc: null pointer");
  SININN_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "wgrad_tc: bad shape");
  SININN_CHECK_ARG(d->taps == 1 || d->taps == 9, "wgrad_tc: taps must be 1 or 9");
  SININN_CHECK_ARG(d->x_dtype == SININN_BF16 && d->dy_dtype == SININN_BF16, "wgrad_tc: operands must be bf16");
  SININN_CHECK_ARG(aligned16(d->x) && aligned16(d->dy) && (d->x_stride % 8) == 0 && (d->dy_stride % 8) == 0,
                   "wgrad_tc: TMA needs 16-byte aligned operands with pixel strides that are multiples of 8 channels "
                   "(x stride %d, dy stride %d)", d->x_stride, d->dy_stride);
  {
    const int rc = launch_wgrad_pair_group(d, 1, d->workspace, d->workspace_bytes, as_stream(stream));
    if (rc != SININN_EUNSUPPORTED) return rc;
  }
  if (d->nterms > 0 || d->nseg > 0) {
    set_error("wgrad_tc: split-operand term lists and merged problems are only taken by the CTA-pair kernel (wide operand > 128 channels)");
    return SININN_EUNSUPPORTED;
  }
  WgradPlan w;
  if (!plan_wgrad(d, w)) {
    set_error("wgrad_tc: unsupported shape (Cin=%d Cout=%d)", d->Cin, d->Cout);
    return SININN_EUNSUPPORTED;
  }
  const size_t need_w = (size_t)w.splits * d->taps * d->Cout * d->Cin * sizeof(float);
  const size_t need = need_w + (d->dbias ? (size_t)2 * w.splits * d->Cout * sizeof(float) : 0);
  if (!d->workspace || d->workspace_bytes < need) {
    set_error("wgrad_tc: workspace too small (%zu < %zu)", d->workspace_bytes, need);
    return SININN_EWORKSPACE;
  }
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("wgrad_tc: cuTensorMapEncodeTiled not available from the driver");
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
    cuuint32_t box[4] = {64, (cuuint32_t)(halo ? WG_HALO_W : WG_BLK), (cuuint32_t)(halo ? WG_HALO_H : WG_BLK), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(which == 0 ? &tmW : &tmN, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("wgrad_tc: cuTensorMapEncodeTiled failed with %d (C=%d stride=%d)", (int)r, C, stride);
      return SININN_ECUDA;
    }
  }
  WgradParams p;
  p.B = d->B; p.H = d->H; p.W = d->W; p.taps = d->taps;
  p.narrow_is_x = w.wide_is_dy;
  p.Cw = w.Cw; p.Cn = w.Cn; p.n_pad = w.n_pad; p.n_groups = w.n_groups;
  p.taps_per_cta = w.taps_per_cta; p.tap_groups = w.tap_groups; p.m_tiles = w.m_tiles; p.splits = w.splits;
  p.blocks_h = w.blocks_h; p.blocks_w = w.blocks_w; p.num_blocks = w.num_blocks; p.blocks_per_split = w.blocks_per_split;
  p.stages = w.stages; p.stage_bytes = w.stage_bytes; p.tx_bytes = w.stage_bytes; p.narrow_bytes = w.narrow_bytes;
  p.partial = reinterpret_cast<float*>(d->workspace);
  p.bias_mode = d->dbias ? (w.wide_is_dy ? 1 : 2) : 0;
  p.bias_partial = d->dbias ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->workspace) + need_w) : nullptr;
  const int bias_rows = d->dbias ? (w.wide_is_dy ? 2 * w.splits : w.splits) : 0;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return SININN_ECUDA;
    }
    attr_set[dev] = true;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + sizeof(Barriers) + 1024;
  const unsigned grid = (unsigned)(w.m_tiles * w.tap_groups * w.splits);
  cudaStream_t st = as_stream(stream);
  launch_k(wgrad_tc_kernel, dim3(grid), dim3(NUM_THREADS), smem, st, tmW, tmN, p);
  SININN_CHECK_LAUNCH("wgrad_tc");
  return launch_reduce_single(d, st, p.partial, w.splits, w.wide_is_dy, p.bias_partial, bias_rows);
}

}  // extern "C"
