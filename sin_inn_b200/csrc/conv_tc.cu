// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate).
//
//   out[p][co] = sum_{tap, ci} in[p + off(tap)][ci] * wpack[tap][co][ci]        (1x1 or 3x3, stride 1, "same")
//
// One CTA tile = 128 output pixels (an 8x16 spatial patch of one image) x up to 256 output channels.
//   warp 0     TMA producer: per K step (tap, 16/32/64-channel slab) one 4-D box load of the shifted
//              activation patch (out-of-bounds rows/columns are zero-filled by TMA = the conv padding) and
//              one 3-D box load of the weight slab, both 128B/64B/32B-swizzled, into a 3..8 stage smem ring
//   warp 1     MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=Cout tile, K=16 per instruction,
//              fp32 accumulator in TMEM (two 256-column buffers => the epilogue of tile i overlaps the
//              MMAs of tile i+1); tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2-9  epilogue, two warps per TMEM lane quarter taking alternate 128-byte output slabs:
//              tcgen05.ld -> + bias, activation, ReLU-bit mask, alpha -> 128B-swizzled smem staging box ->
//              TMA tensor store (or TMA reduce-add for fp32 accumulation) with hardware clipping of partial
//              tiles; optionally emits the ReLU sign bits of its output (1 bit / element) for the backward pass
// The grid is persistent (one CTA per SM, static round-robin over tiles).
//
// Replaces cuDNN's nn.Conv2d forward / data-gradient inside subnet_conv / subnet_conv_1x1
// (/root/reference/archs.py:11-17) and DenseBlock (/root/reference/archs.py:77-81,88-95).
#include "tc_common.cuh"

namespace sininn {
namespace tc {

constexpr int TILE_H = 8, TILE_W = 16, TILE_M = TILE_H * TILE_W;   // 128 pixels = UMMA M
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int ACC_STRIDE = 256;          // columns between the two accumulator buffers
constexpr int STAGING_BYTES = 32 * 128;  // one epilogue warp: 32 rows x 128 B
constexpr int EPI_SMEM_BYTES = 256 * 4 + NUM_EPI_WARPS * STAGING_BYTES;
constexpr int SMEM_RING_BUDGET = 227 * 1024 - EPI_SMEM_BYTES - BARRIER_BYTES - 1024;   // 1024: worst-case alignment pad

struct Params {
  int B, H, W, Cin, Cout;
  int taps, kc, k_chunks;        // kc = channels per K step (16/32/64), k_chunks = ceil(Cin / kc)
  int n_tile, n_tiles;           // output channels per CTA tile (multiple of 16, <= 256), tiles along N
  int tiles_h, tiles_w;
  long long num_tiles;           // B * tiles_h * tiles_w * n_tiles
  int stages;
  uint32_t a_bytes, b_bytes;     // smem bytes per stage (each a multiple of 1024)
  uint32_t tx_bytes;             // bytes TMA delivers per stage (A box + B box)
  uint32_t sbo;                  // 8 rows * row bytes
  uint32_t layout_type;          // UMMA smem-descriptor swizzle code
  const float* bias;
  void* out; int out_f32; int out_stride;
  int act; float slope;
  const void* mask; int mask_stride; int mask_act;     // element mask (fallback path only)
  const uint32_t* bits_in;       // ReLU sign bits of the activation this gradient flows through, [npix][bit_words]
  uint32_t* bits_out;            // sign bits of this kernel's own output, [npix][bit_words]
  int bit_words;
  int accumulate; float alpha;
  int tma_out;                   // 1: TMA-store epilogue; 0: per-thread fallback (unaligned output slices)
};

__device__ __forceinline__ void tile_coords(const Params& p, long long t, int& b, int& h0, int& w0, int& n0) {
  int nt = (int)(t % p.n_tiles); t /= p.n_tiles;
  int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
  int th = (int)(t % p.tiles_h);
  b = (int)(t / p.tiles_h);
  h0 = th * TILE_H; w0 = tw * TILE_W; n0 = nt * p.n_tile;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// per-thread fallback for output slices TMA cannot address (unaligned base / stride)
template <typename TO>
__device__ __forceinline__ void epilogue_chunk(const Params& p, const uint32_t (&v)[16], long long pix, int col0, bool row_ok) {
  if (!row_ok) return;
  TO* __restrict__ out = reinterpret_cast<TO*>(p.out) + pix * p.out_stride + col0;
  const TO* __restrict__ mask = p.mask ? reinterpret_cast<const TO*>(p.mask) + pix * p.mask_stride + col0 : nullptr;
  const int ncol = min(16, p.Cout - col0);
  for (int j = 0; j < ncol; ++j) {
    float x = __uint_as_float(v[j]);
    if (p.bias != nullptr) x += __ldg(p.bias + col0 + j);
    x = act_fwd(p.act, p.slope, x);
    if (mask != nullptr) x *= act_grad(p.mask_act, p.slope, to_f32(mask[j]));
    if (p.bits_in != nullptr) {
      const uint32_t w = p.bits_in[pix * p.bit_words + ((col0 + j) >> 5)];
      if (!((w >> ((col0 + j) & 31)) & 1u)) x = 0.f;
    }
    x *= p.alpha;
    if (p.accumulate) x += to_f32(out[j]);
    out[j] = from_f32<TO>(x);
  }
}

// One 128-byte output slab (64 bf16 or 32 fp32 columns) of a warp's 32 accumulator rows:
// registers -> (+bias, activation, sign-bit mask, alpha) -> 128B-swizzled staging rows in shared memory.
// v holds the slab's accumulator columns; returns the sign bits of the produced values (bit j = value j > 0).
template <int NCOL>
__device__ __forceinline__ void slab_math(const Params& p, float (&x)[NCOL], const float* bias_s, const uint32_t* mbits) {
#pragma unroll
  for (int q = 0; q < NCOL / 4; ++q) {
    const float4 bq = *reinterpret_cast<const float4*>(bias_s + 4 * q);
    x[4 * q + 0] = act_fwd(p.act, p.slope, x[4 * q + 0] + bq.x);
    x[4 * q + 1] = act_fwd(p.act, p.slope, x[4 * q + 1] + bq.y);
    x[4 * q + 2] = act_fwd(p.act, p.slope, x[4 * q + 2] + bq.z);
    x[4 * q + 3] = act_fwd(p.act, p.slope, x[4 * q + 3] + bq.w);
  }
  if (mbits != nullptr) {
#pragma unroll
    for (int j = 0; j < NCOL; ++j)
      if (!((mbits[j >> 5] >> (j & 31)) & 1u)) x[j] = 0.f;
  }
  if (p.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < NCOL; ++j) x[j] *= p.alpha;
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // every pointer is smem_raw + offset so the compiler keeps the shared address space (LDS/STS, not generic)
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* ring = smem_raw + pad;
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;                      // both multiples of 1024
  const uint32_t ring_bytes = (uint32_t)p.stages * stage_bytes;
  uint8_t* staging = ring + ring_bytes;                                    // [8 warps][32 rows][128 B], 1024-aligned
  Barriers* bars = reinterpret_cast<Barriers*>(staging + NUM_EPI_WARPS * STAGING_BYTES);
  float* bias_s = reinterpret_cast<float*>(staging + NUM_EPI_WARPS * STAGING_BYTES + BARRIER_BYTES);   // [256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring_u32 = smem_u32(ring);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bars->acc_full[a]), 1);
      mbar_init(smem_u32(&bars->acc_empty[a]), NUM_EPI_WARPS);     // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmO);
  }
  if (warp == 1) {   // TMEM allocation (whole warp), address lands in smem
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int k_steps = p.taps * p.k_chunks;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int b, h0, w0, n0;
        tile_coords(p, t, b, h0, w0, n0);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
            const uint32_t full = smem_u32(&bars->full[stage]);
            mbar_expect_tx(full, p.tx_bytes);
            const uint32_t a_dst = ring_u32 + stage * stage_bytes;
            tma_load_4d(a_dst, &tmA, full, kc * p.kc, w0 + dx, h0 + dy, b);
            tma_load_3d(a_dst + p.a_bytes, &tmB, full, kc * p.kc, n0, tap);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const int mma_per_step = p.kc / 16;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(smem_u32(&bars->acc_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int ks = 0; ks < k_steps; ++ks) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = ring_u32 + stage * stage_bytes;
          const uint32_t b_addr = a_addr + p.a_bytes;
          for (int k = 0; k < mma_per_step; ++k) {
            const uint64_t adesc = make_desc(a_addr + k * 32, p.sbo, p.layout_type);
            const uint64_t bdesc = make_desc(b_addr + k * 32, p.sbo, p.layout_type);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bars->empty[stage]));        // frees the smem stage when these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&bars->acc_full[acc]));          // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ======================= epilogue (warps 2..9) =======================
    const int ew = warp - 2;                                   // 0..7
    const int quarter = warp & 3;                              // TMEM lanes 32*quarter .. +31 (hardware rule: warp id % 4)
    const int half = ew >> 2;                                  // which of the two warps of this quarter
    const int row = quarter * 32 + lane;                       // pixel row inside the tile
    const int hl = row / TILE_W, wl = row % TILE_W;
    uint8_t* stg = staging + ew * STAGING_BYTES;
    const uint32_t stg_u32 = smem_u32(stg);
    const int etid = threadIdx.x - 64;                         // 0..255 among the epilogue threads
    const int slab_cols = p.out_f32 ? 32 : 64;                 // 128 bytes of output per row
    int acc = 0; uint32_t acc_phase = 0;
    int bias_n0 = -1;
    for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int b, h0, w0, n0;
      tile_coords(p, t, b, h0, w0, n0);
      const int oh = h0 + hl, ow = w0 + wl;
      const bool row_ok = (oh < p.H) && (ow < p.W);
      const long long pix = ((long long)b * p.H + oh) * p.W + ow;
      if (n0 != bias_n0) {                                     // (re)load the bias slice of this N tile
        asm volatile("bar.sync 1, 256;" ::: "memory");         // everyone done with the previous slice
        {
          const int co = n0 + etid;
          bias_s[etid] = (p.bias != nullptr && co < p.Cout) ? __ldg(p.bias + co) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        bias_n0 = n0;
      }
      mbar_wait(smem_u32(&bars->acc_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE;
      const int n_valid = min(p.n_tile, p.Cout - n0);
      if (p.tma_out) {
        const int n_slabs = (n_valid + slab_cols - 1) / slab_cols;
        for (int s = half; s < n_slabs; s += 2) {
          const int c = s * slab_cols;                         // first accumulator column of the slab
          // sign-bit mask words of this row for the slab's columns
          uint32_t mb[2] = {0xffffffffu, 0xffffffffu};
          if (p.bits_in != nullptr) {
            const int w0i = (n0 + c) >> 5;
            mb[0] = row_ok ? __ldg(p.bits_in + pix * p.bit_words + w0i) : 0u;
            if (!p.out_f32) mb[1] = (row_ok && w0i + 1 < p.bit_words) ? __ldg(p.bits_in + pix * p.bit_words + w0i + 1) : 0u;
          }
          if (lane == 0) bulk_wait_read0();                    // previous TMA store has finished reading the staging rows
          __syncwarp();
          uint32_t sign[2] = {0u, 0u};
          if (p.out_f32) {
            uint32_t v[32];
            tmem_ld32(t_base + c, v);
            tmem_ld_wait();
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
            slab_math<32>(p, x, bias_s + c, p.bits_in ? mb : nullptr);
#pragma unroll
            for (int j = 0; j < 32; ++j) sign[0] |= (x[j] > 0.f ? 1u : 0u) << j;
#pragma unroll
            for (int q = 0; q < 8; ++q)                         // 16-byte piece q of the row, 128B swizzle: q ^ (row & 7)
              *reinterpret_cast<float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                  make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
          } else {
            uint32_t v0[32], v1[32];
            tmem_ld32(t_base + c, v0);
            tmem_ld32(t_base + c + 32, v1);                    // (columns past n_valid are clipped by the TMA store)
            tmem_ld_wait();
            float x[64];
#pragma unroll
            for (int j = 0; j < 32; ++j) { x[j] = __uint_as_float(v0[j]); x[32 + j] = __uint_as_float(v1[j]); }
            slab_math<64>(p, x, bias_s + c, p.bits_in ? mb : nullptr);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              sign[0] |= (x[j] > 0.f ? 1u : 0u) << j;
              sign[1] |= (x[32 + j] > 0.f ? 1u : 0u) << j;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint4 o;
              o.x = pack_bf16(x[8 * q + 0], x[8 * q + 1]);
              o.y = pack_bf16(x[8 * q + 2], x[8 * q + 3]);
              o.z = pack_bf16(x[8 * q + 4], x[8 * q + 5]);
              o.w = pack_bf16(x[8 * q + 6], x[8 * q + 7]);
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
            }
          }
          if (p.bits_out != nullptr && row_ok) {
            const int w0i = (n0 + c) >> 5;
            p.bits_out[pix * p.bit_words + w0i] = sign[0];
            if (!p.out_f32 && w0i + 1 < p.bit_words) p.bits_out[pix * p.bit_words + w0i + 1] = sign[1];
          }
          fence_async_smem();                                  // generic-proxy smem writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            // this warp's 32 rows are tile rows h = 2*quarter, 2*quarter+1 (16 pixels each): box {128 B, 16, 2, 1}
            if (p.accumulate) tma_reduce_add_4d(&tmO, stg_u32, n0 + c, w0, h0 + 2 * quarter, b);
            else tma_store_4d(&tmO, stg_u32, n0 + c, w0, h0 + 2 * quarter, b);
            bulk_commit();
          }
        }
      } else if (half == 0) {
        for (int c = 0; c < n_valid; c += 16) {
          uint32_t v[16];
          tmem_ld16(t_base + c, v);
          tmem_ld_wait();
          if (p.out_f32) epilogue_chunk<float>(p, v, pix, n0 + c, row_ok);
          else epilogue_chunk<__nv_bfloat16>(p, v, pix, n0 + c, row_ok);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_out && lane == 0) bulk_wait_all();               // outstanding tensor stores complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace sininn

using namespace sininn;

extern "C" {

int sininn_conv_tc(const sininn_conv_desc* d, sininn_stream_t stream) {
  using namespace sininn::tc;
  SININN_CHECK_ARG(d != nullptr && d->in && d->wpack && d->out, "conv_tc: null pointer");
  SININN_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv_tc: bad shape");
  SININN_CHECK_ARG(d->taps == 1 || d->taps == 9, "conv_tc: taps must be 1 or 9");
  SININN_CHECK_ARG(d->in_dtype == SININN_BF16, "conv_tc: operands must be bf16");
  SININN_CHECK_ARG(d->out_dtype == SININN_BF16 || d->out_dtype == SININN_F32, "conv_tc: bad out_dtype");
  SININN_CHECK_ARG((d->k_pad % 16) == 0 && (d->rows_pad % 16) == 0 && d->k_pad >= d->Cin && d->rows_pad >= d->Cout,
                   "conv_tc: packed weights must be padded to multiples of 16 (rows_pad=%d k_pad=%d)", d->rows_pad, d->k_pad);
  SININN_CHECK_ARG(aligned16(d->in) && ((long long)d->in_stride * 2) % 16 == 0,
                   "conv_tc: TMA needs a 16-byte aligned input slice and a pixel stride that is a multiple of 8 channels "
                   "(stride %d)", d->in_stride);
  SININN_CHECK_ARG(aligned16(d->wpack), "conv_tc: packed weights misaligned");
  SININN_CHECK_ARG(!(d->mask && d->mask_bits), "conv_tc: give either an element mask or a bit mask");
  SININN_CHECK_ARG(!(d->accumulate && d->out_dtype != SININN_F32), "conv_tc: accumulation needs an fp32 output");
  const int bit_words = (d->Cout + 31) / 32;
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("conv_tc: cuTensorMapEncodeTiled not available from the driver");
    return SININN_ECUDA;
  }
  Params p;
  p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.taps = d->taps;
  p.kc = pick_kc(d->k_pad);
  p.k_chunks = (d->Cin + p.kc - 1) / p.kc;      // slabs beyond Cin would be all zero: skipped
  p.n_tile = d->rows_pad <= 256 ? d->rows_pad : 256;
  p.n_tiles = (d->rows_pad + p.n_tile - 1) / p.n_tile;
  p.tiles_h = (d->H + TILE_H - 1) / TILE_H;
  p.tiles_w = (d->W + TILE_W - 1) / TILE_W;
  p.num_tiles = (long long)d->B * p.tiles_h * p.tiles_w * p.n_tiles;
  const uint32_t row_bytes = (uint32_t)p.kc * 2;
  p.a_bytes = TILE_M * row_bytes;                                  // 4 / 8 / 16 KiB
  p.b_bytes = ((uint32_t)p.n_tile * row_bytes + 1023u) & ~1023u;   // keep every stage 1024-byte aligned
  p.tx_bytes = p.a_bytes + (uint32_t)p.n_tile * row_bytes;
  p.sbo = 8 * row_bytes;
  p.layout_type = p.kc == 64 ? 2u : (p.kc == 32 ? 4u : 6u);
  int stages = SMEM_RING_BUDGET / (int)(p.a_bytes + p.b_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  SININN_CHECK_ARG(stages >= 2, "conv_tc: tile does not fit in shared memory");
  p.stages = stages;
  p.bias = d->bias;
  p.out = d->out; p.out_f32 = d->out_dtype == SININN_F32; p.out_stride = d->out_stride;
  p.act = d->act; p.slope = d->slope;
  p.mask = d->mask; p.mask_stride = d->mask_stride; p.mask_act = d->mask_act;
  p.bits_in = reinterpret_cast<const uint32_t*>(d->mask_bits);
  p.bits_out = reinterpret_cast<uint32_t*>(d->bits_out);
  p.bit_words = bit_words;
  p.accumulate = d->accumulate; p.alpha = d->alpha;
  const int esz = p.out_f32 ? 4 : 2;
  // TMA epilogue needs a 16-byte aligned output slice / pixel stride and no per-element mask tensor
  p.tma_out = (aligned16(d->out) && ((long long)d->out_stride * esz) % 16 == 0 && d->mask == nullptr) ? 1 : 0;

  const CUtensorMapSwizzle swz = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB, tmO;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_stride * 2, (cuuint64_t)d->W * d->in_stride * 2,
                             (cuuint64_t)d->H * d->W * d->in_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.kc, TILE_W, TILE_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_tc: cuTensorMapEncodeTiled(activations) failed with %d (Cin=%d stride=%d W=%d H=%d B=%d kc=%d)", (int)r,
                d->Cin, d->in_stride, d->W, d->H, d->B, p.kc);
      return SININN_ECUDA;
    }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->k_pad, (cuuint64_t)d->rows_pad, (cuuint64_t)d->taps};
    cuuint64_t strides[2] = {(cuuint64_t)d->k_pad * 2, (cuuint64_t)d->rows_pad * d->k_pad * 2};
    cuuint32_t box[3] = {(cuuint32_t)p.kc, (cuuint32_t)p.n_tile, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->wpack), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_tc: cuTensorMapEncodeTiled(weights) failed with %d (k_pad=%d rows_pad=%d taps=%d)", (int)r, d->k_pad,
                d->rows_pad, d->taps);
      return SININN_ECUDA;
    }
  }
  if (p.tma_out) {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->out_stride * esz, (cuuint64_t)d->W * d->out_stride * esz,
                             (cuuint64_t)d->H * d->W * d->out_stride * esz};
    cuuint32_t box[4] = {(cuuint32_t)(128 / esz), TILE_W, 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmO, p.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d->out, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv_tc: cuTensorMapEncodeTiled(output) failed with %d (Cout=%d stride=%d)", (int)r, d->Cout, d->out_stride);
      return SININN_ECUDA;
    }
  } else {
    tmO = tmA;   // unused
  }
  const size_t smem = (size_t)p.stages * (p.a_bytes + p.b_bytes) + EPI_SMEM_BYTES + BARRIER_BYTES + 1024;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return SININN_ECUDA;
    }
    attr_set[dev] = true;
  }
  long long grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  conv_tc_kernel<<<(unsigned)grid, NUM_THREADS, smem, as_stream(stream)>>>(tmA, tmB, tmO, p);
  SININN_CHECK_LAUNCH("conv_tc");
  return SININN_OK;
}

}  // extern "C"
