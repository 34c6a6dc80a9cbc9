// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate).
//
//   out[p][co] = sum_{tap, ci} in[p + off(tap)][ci] * wpack[tap][co][ci]        (1x1 or 3x3, stride 1, "same")
//
// One CTA tile = 128 output pixels (an 8x16 spatial patch of one image) x up to 256 output channels.
//   warp 0   TMA producer: per K step (tap, 16/32/64-channel slab) one 4-D box load of the shifted
//            activation patch (out-of-bounds rows/columns are zero-filled by TMA = the conv padding) and
//            one 3-D box load of the weight slab, both 128B/64B/32B-swizzled, into a 3..8 stage smem ring
//   warp 1   MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=Cout tile, K=16 per instruction,
//            fp32 accumulator in TMEM (two 256-column buffers => the epilogue of tile i overlaps the
//            MMAs of tile i+1); tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 16 columns), bias / activation / activation-derivative
//            mask / alpha / accumulate, 128-bit global stores (bf16 or fp32, strided channel slices)
// The grid is persistent (one CTA per SM, static round-robin over tiles).
//
// Replaces cuDNN's nn.Conv2d forward / data-gradient inside subnet_conv / subnet_conv_1x1
// (/root/reference/archs.py:11-17) and DenseBlock (/root/reference/archs.py:77-81,88-95).
#include <cuda.h>
#include "common.cuh"

namespace sininn {
int wgrad_simt_splits(const sininn_wgrad_desc* d);

namespace tc {

constexpr int TILE_H = 8, TILE_W = 16, TILE_M = TILE_H * TILE_W;   // 128 pixels = UMMA M
constexpr int MAX_STAGES = 8;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;          // columns between the two accumulator buffers
constexpr int SMEM_RING_BUDGET = 200 * 1024;

struct Params {
  int B, H, W, Cin, Cout;
  int taps, kc, k_chunks;        // kc = channels per K step (16/32/64), k_chunks = ceil(k_pad / kc)
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
  const void* mask; int mask_stride; int mask_act;
  int accumulate; float alpha;
  int vec_out;                   // rows may be written/read 16 B at a time
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand with 32/64/128-byte swizzle:
// start>>4 | LBO>>4 (unused for swizzled K-major, canonical 1) | SBO>>4 | version 1 (bit 46) | layout (bits 61..63)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

struct __align__(8) Barriers {
  uint64_t full[MAX_STAGES];
  uint64_t empty[MAX_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void tile_coords(const Params& p, long long t, int& b, int& h0, int& w0, int& n0) {
  int nt = (int)(t % p.n_tiles); t /= p.n_tiles;
  int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
  int th = (int)(t % p.tiles_h);
  b = (int)(t / p.tiles_h);
  h0 = th * TILE_H; w0 = tw * TILE_W; n0 = nt * p.n_tile;
}

template <typename TO>
__device__ __forceinline__ void epilogue_chunk(const Params& p, const uint32_t (&v)[16], long long pix, int col0, bool row_ok) {
  if (!row_ok) return;
  TO* __restrict__ out = reinterpret_cast<TO*>(p.out) + pix * p.out_stride + col0;
  const TO* __restrict__ mask = p.mask ? reinterpret_cast<const TO*>(p.mask) + pix * p.mask_stride + col0 : nullptr;
  const int ncol = min(16, p.Cout - col0);
  float r[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x = __uint_as_float(v[j]);
    if (p.bias != nullptr && j < ncol) x += __ldg(p.bias + col0 + j);
    r[j] = act_fwd(p.act, p.slope, x);
  }
  if (ncol == 16 && p.vec_out) {
    if (mask != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 m = load4(mask + 4 * q);
        r[4 * q + 0] *= act_grad(p.mask_act, p.slope, m.x);
        r[4 * q + 1] *= act_grad(p.mask_act, p.slope, m.y);
        r[4 * q + 2] *= act_grad(p.mask_act, p.slope, m.z);
        r[4 * q + 3] *= act_grad(p.mask_act, p.slope, m.w);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 o = make_float4(r[4 * q] * p.alpha, r[4 * q + 1] * p.alpha, r[4 * q + 2] * p.alpha, r[4 * q + 3] * p.alpha);
      if (p.accumulate) {
        float4 old = load4(out + 4 * q);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      store4(out + 4 * q, o);
    }
  } else {
    for (int j = 0; j < ncol; ++j) {
      float x = r[j];
      if (mask != nullptr) x *= act_grad(p.mask_act, p.slope, to_f32(mask[j]));
      x *= p.alpha;
      if (p.accumulate) x += to_f32(out[j]);
      out[j] = from_f32<TO>(x);
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // ring first (1024-byte aligned for the swizzle atoms), barriers behind it
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;       // both multiples of 1024
  Barriers* bars = reinterpret_cast<Barriers*>(ring + (size_t)p.stages * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring_u32 = smem_u32(ring);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bars->acc_full[a]), 1);
      mbar_init(smem_u32(&bars->acc_empty[a]), 4);           // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
    // ======================= epilogue (warps 2..5) =======================
    const int quarter = warp & 3;                              // TMEM lanes 32*quarter .. +31
    const int row = quarter * 32 + lane;                       // pixel row inside the tile
    const int hl = row / TILE_W, wl = row % TILE_W;
    int acc = 0; uint32_t acc_phase = 0;
    for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int b, h0, w0, n0;
      tile_coords(p, t, b, h0, w0, n0);
      const int oh = h0 + hl, ow = w0 + wl;
      const bool row_ok = (oh < p.H) && (ow < p.W);
      const long long pix = ((long long)b * p.H + oh) * p.W + ow;
      mbar_wait(smem_u32(&bars->acc_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE;
      const int n_valid = min(p.n_tile, p.Cout - n0);
      for (int c = 0; c < n_valid; c += 16) {
        uint32_t v[16];
        tmem_ld16(t_base + c, v);
        tmem_ld_wait();
        if (p.out_f32) epilogue_chunk<float>(p, v, pix, n0 + c, row_ok);
        else epilogue_chunk<__nv_bfloat16>(p, v, pix, n0 + c, row_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static int pick_kc(int k_pad) {
  if (k_pad % 64 == 0) return 64;
  if (k_pad % 32 == 0) return 32;
  return 16;
}

}  // namespace tc
}  // namespace sininn

using namespace sininn;

extern "C" {

size_t sininn_wgrad_workspace_bytes(const sininn_wgrad_desc* d, int tensor_core) {
  if (!d || d->Cin <= 0 || d->Cout <= 0 || d->taps <= 0) return 0;
  (void)tensor_core;
  return (size_t)wgrad_simt_splits(d) * d->taps * d->Cout * d->Cin * sizeof(float);
}

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
  p.accumulate = d->accumulate; p.alpha = d->alpha;
  const int esz = p.out_f32 ? 4 : 2;
  bool vec = aligned16(d->out) && ((long long)d->out_stride * esz) % 16 == 0;
  if (d->mask) vec = vec && aligned16(d->mask) && ((long long)d->mask_stride * esz) % 16 == 0;
  p.vec_out = vec ? 1 : 0;

  p.tx_bytes = p.a_bytes + (uint32_t)p.n_tile * row_bytes;
  const CUtensorMapSwizzle swz = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmA, tmB;
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
  const size_t smem = (size_t)p.stages * (p.a_bytes + p.b_bytes) + sizeof(Barriers) + 1024;
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
  conv_tc_kernel<<<(unsigned)grid, NUM_THREADS, smem, as_stream(stream)>>>(tmA, tmB, p);
  SININN_CHECK_LAUNCH("conv_tc");
  return SININN_OK;
}

int sininn_wgrad_tc(const sininn_wgrad_desc* d, sininn_stream_t stream) {
  (void)d; (void)stream;
  set_error("wgrad_tc: not built yet");
  return SININN_EUNSUPPORTED;
}

}  // extern "C"
