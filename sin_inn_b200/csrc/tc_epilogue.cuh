// Shared pieces of the tcgen05 convolution kernels (conv_tc.cu, conv_tc_halo.cu): parameter block and the
// TMA-store epilogue (TMEM -> registers -> 128B-swizzled smem staging -> tensor store / reduce-add).
#pragma once
#include "tc_common.cuh"

namespace sininn {
namespace tc {

constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int ACC_STRIDE = 256;          // columns between the two accumulator buffers
constexpr int STAGING_BYTES = 32 * 128;  // one epilogue warp: 32 rows x 128 B
constexpr int EPI_SMEM_BYTES = 256 * 4 + NUM_EPI_WARPS * STAGING_BYTES;
constexpr int SMEM_RING_BUDGET = 227 * 1024 - EPI_SMEM_BYTES - BARRIER_BYTES - 1024;   // 1024: worst-case alignment pad

// affine coupling fused into an epilogue (sininn_conv_desc / sininn_subnet1x1_desc cpl_*): mode 0 off, 1 apply, 2 backward
struct CplParams {
  int mode, L, inverse;
  float clamp;
  float* u; int u_stride;
  float* du; int du_stride;
  __nv_bfloat16* bf16;
  __nv_bfloat16* da;
  float* a;                      // mode 1: optional copy of the subnet output [s | t], fp32 [npix][2L]
};

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
  CplParams cpl;                 // conv_tc_pair only
};

template <int TW>
__device__ __forceinline__ void tile_coords(const Params& p, long long t, int& b, int& h0, int& w0, int& n0) {
  constexpr int TH = 128 / TW;
  int nt = (int)(t % p.n_tiles); t /= p.n_tiles;
  int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
  int th = (int)(t % p.tiles_h);
  b = (int)(t / p.tiles_h);
  h0 = th * TH; w0 = tw * TW; n0 = nt * p.n_tile;
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

// ReLU sign-bit masks (bits_out / mask_bits of the C ABI): one bit per element, 32 elements per word.  Element e of a word
// sits at bit (e >> 1) + 16 * (e & 1): the two halves of the packed bf16x2 register j of a word (elements 2j, 2j + 1) map to
// bits j and 16 + j, so producing a word costs one HSET2 + one LOP3 per REGISTER and applying it one shift-AND, one multiply
// (FMA pipe) and one AND per register.  The epilogues are bound by the ALU pipe (16 lanes per clock and scheduler for
// FMNMX / FSETP / SEL / LOP3 / F2FP -- measured with in-kernel stamps), not by issue slots or TMEM reads.
__host__ __device__ constexpr int sign_bit_pos(int e) { return (e >> 1) + 16 * (e & 1); }
// max(v, 0) on both bf16 halves (relu(round(x)) == round(relu(x)): rounding keeps the sign)
__device__ __forceinline__ uint32_t bf16x2_relu(uint32_t v) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&v), __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// 0xFFFF in each half of the result whose bf16 half of v is > 0
__device__ __forceinline__ uint32_t bf16x2_gt0_mask(uint32_t v) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v), __floats2bfloat162_rn(0.f, 0.f));
}
// the AND mask of packed register j (0..15) of a sign-bit word: 0xFFFF per half whose bit is set
__device__ __forceinline__ uint32_t sign_bits_expand(uint32_t word, int j) { return ((word >> j) & 0x00010001u) * 0xFFFFu; }

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
      if (!((w >> sign_bit_pos((col0 + j) & 31)) & 1u)) x = 0.f;
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
  // The epilogue warps are instruction-bound on the 256-wide outputs (measured: ~1600 cycles per 64-column slab and
  // warp), so every variant is branch-free per element and does only what its launch needs; the branches below are
  // warp-uniform and outside the element loops.
  if (p.bias != nullptr || p.act != SININN_ACT_NONE) {
    if (p.act == SININN_ACT_RELU) {                      // FADD + FMNMX per element
#pragma unroll
      for (int q = 0; q < NCOL / 4; ++q) {
        const float4 bq = *reinterpret_cast<const float4*>(bias_s + 4 * q);
        x[4 * q + 0] = fmaxf(x[4 * q + 0] + bq.x, 0.f);
        x[4 * q + 1] = fmaxf(x[4 * q + 1] + bq.y, 0.f);
        x[4 * q + 2] = fmaxf(x[4 * q + 2] + bq.z, 0.f);
        x[4 * q + 3] = fmaxf(x[4 * q + 3] + bq.w, 0.f);
      }
    } else {
      // f(v) = max(v, ns * v): identity for ns = 1 (no activation), LeakyReLU for 0 <= ns = slope <= 1
      const float ns = p.act == SININN_ACT_LRELU ? p.slope : 1.f;
#pragma unroll
      for (int q = 0; q < NCOL / 4; ++q) {
        const float4 bq = *reinterpret_cast<const float4*>(bias_s + 4 * q);
        const float b4[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = x[4 * q + e] + b4[e];
          x[4 * q + e] = ns == 1.f ? v : (v > 0.f ? v : ns * v);
        }
      }
    }
  }
  if (mbits != nullptr) {
#pragma unroll
    for (int j = 0; j < NCOL; ++j)
      if (!((mbits[j >> 5] >> sign_bit_pos(j & 31)) & 1u)) x[j] = 0.f;
  }
  if (p.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < NCOL; ++j) x[j] *= p.alpha;
  }
}

// (Re)load the bias slice of N tile n0 into bias_s (all 256 epilogue threads take part; named barrier 1).
__device__ __forceinline__ void epilogue_load_bias(const Params& p, float* bias_s, int n0, int etid, int& bias_n0) {
  if (n0 == bias_n0) return;
  asm volatile("bar.sync 1, 256;" ::: "memory");           // everyone done with the previous slice
  {
    int co = n0 + etid;
    const bool ok = p.bias != nullptr && co < p.Cout;
    if (p.cpl.mode != 0) co = (co & 1) ? p.cpl.L + (co >> 1) : (co >> 1);     // rows interleaved (s_0, t_0, s_1, t_1, ...)
    bias_s[etid] = ok ? __ldg(p.bias + co) : 0.f;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  bias_n0 = n0;
}

// One output tile of one epilogue warp: the warp's 32 accumulator rows (TMEM lanes 32*quarter..+31, columns from
// t_base) are 32/TW consecutive tile rows, stored as one TMA box {128 B, TW, 32/TW, 1} per 128-byte output slab.
// The two warps of a lane quarter (half = 0/1) take alternate slabs.  Tile origin (b, h0, w0), first channel n0;
// a tile with b >= p.B is a phantom (nothing is stored).
#ifdef SININN_PAIR_TRACE
#define EPI_STAMP() do { if (etrace != nullptr && lane == 0) { *etrace++ = clock64(); } } while (0)
#else
#define EPI_STAMP() do { } while (0)
#endif
template <int TW>
__device__ __forceinline__ void epilogue_tile(const Params& p, const CUtensorMap* tmO, uint8_t* stg, const float* bias_s,
                                              uint32_t t_base, int b, int h0, int w0, int n0, int quarter, int half, int lane,
                                              long long* etrace = nullptr) {
  const uint32_t stg_u32 = smem_u32(stg);
  const int row = quarter * 32 + lane;                       // pixel row inside the tile
  const int hl = row / TW, wl = row % TW;
  const int slab_cols = p.out_f32 ? 32 : 64;                 // 128 bytes of output per row
  const int oh = h0 + hl, ow = w0 + wl;
  const bool row_ok = (oh < p.H) && (ow < p.W) && (b < p.B);
  const long long pix = ((long long)b * p.H + oh) * p.W + ow;
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
      EPI_STAMP();
      if (lane == 0) bulk_wait_read0();                    // previous TMA store has finished reading the staging rows
      __syncwarp();
      EPI_STAMP();
      uint32_t sign[2] = {0u, 0u};
      if (p.out_f32) {
        uint32_t v[32];
        tmem_ld32(t_base + c, v);
        tmem_ld_wait();
        EPI_STAMP();
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        slab_math<32>(p, x, bias_s + c, p.bits_in ? mb : nullptr);
        if (p.bits_out != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sign[0] |= (x[j] > 0.f ? 1u : 0u) << sign_bit_pos(j);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)                         // 16-byte piece q of the row, 128B swizzle: q ^ (row & 7)
          *reinterpret_cast<float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
              make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
      } else {
        uint32_t v0[32], v1[32];
        tmem_ld32(t_base + c, v0);
        tmem_ld32(t_base + c + 32, v1);                    // (columns past n_valid are clipped by the TMA store)
        tmem_ld_wait();
        EPI_STAMP();
        // the two shapes that dominate a training step work on packed bf16x2 registers (half the ALU-pipe instructions):
        //   conv1 of a subnet:            bias + ReLU (+ sign bits for the backward pass)
        //   masked data gradient (conv2): nothing but the stored sign-bit mask
        const bool fast_relu = p.act == SININN_ACT_RELU && p.bits_in == nullptr && p.alpha == 1.0f;
        const bool fast_mask = p.act == SININN_ACT_NONE && p.bias == nullptr && p.bits_in != nullptr && p.bits_out == nullptr && p.alpha == 1.0f;
        if (fast_relu || fast_mask) {
          uint32_t pk[32];
          if (fast_relu) {
            const float* bsl = bias_s + c;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 ba = *reinterpret_cast<const float4*>(bsl + 4 * q), bb = *reinterpret_cast<const float4*>(bsl + 32 + 4 * q);
              pk[2 * q] = bf16x2_relu(pack_bf16(__uint_as_float(v0[4 * q]) + ba.x, __uint_as_float(v0[4 * q + 1]) + ba.y));
              pk[2 * q + 1] = bf16x2_relu(pack_bf16(__uint_as_float(v0[4 * q + 2]) + ba.z, __uint_as_float(v0[4 * q + 3]) + ba.w));
              pk[16 + 2 * q] = bf16x2_relu(pack_bf16(__uint_as_float(v1[4 * q]) + bb.x, __uint_as_float(v1[4 * q + 1]) + bb.y));
              pk[16 + 2 * q + 1] = bf16x2_relu(pack_bf16(__uint_as_float(v1[4 * q + 2]) + bb.z, __uint_as_float(v1[4 * q + 3]) + bb.w));
            }
            if (p.bits_out != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                sign[0] |= bf16x2_gt0_mask(pk[j]) & ((1u << j) | (1u << (16 + j)));
                sign[1] |= bf16x2_gt0_mask(pk[16 + j]) & ((1u << j) | (1u << (16 + j)));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              pk[j] = pack_bf16(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1])) & sign_bits_expand(mb[0], j);
              pk[16 + j] = pack_bf16(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])) & sign_bits_expand(mb[1], j);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        } else {
          float x[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) { x[j] = __uint_as_float(v0[j]); x[32 + j] = __uint_as_float(v1[j]); }
          slab_math<64>(p, x, bias_s + c, p.bits_in ? mb : nullptr);
          if (p.bits_out != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              sign[0] |= (x[j] > 0.f ? 1u : 0u) << sign_bit_pos(j);
              sign[1] |= (x[32 + j] > 0.f ? 1u : 0u) << sign_bit_pos(j);
            }
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
      }
      EPI_STAMP();
      if (p.bits_out != nullptr && row_ok) {
        const int w0i = (n0 + c) >> 5;
        p.bits_out[pix * p.bit_words + w0i] = sign[0];
        if (!p.out_f32 && w0i + 1 < p.bit_words) p.bits_out[pix * p.bit_words + w0i + 1] = sign[1];
      }
      fence_async_smem();                                  // generic-proxy smem writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        // this warp's 32 rows are 32/TW consecutive tile rows: box {128 B, TW, 32/TW, 1}
        if (p.accumulate) tma_reduce_add_4d(tmO, stg_u32, n0 + c, w0, h0 + (32 / TW) * quarter, b);
        else tma_store_4d(tmO, stg_u32, n0 + c, w0, h0 + (32 / TW) * quarter, b);
        bulk_commit();
      }
      EPI_STAMP();
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
}

// Affine coupling in the epilogue (GLOW, archs.py:61-64; equations SURVEY.md 8a).  The accumulator columns are the
// interleaved subnet output (s_0, t_0, s_1, t_1, ...): one 32-column slab = 16 channels of this thread's pixel.
//   cpl_mode 1:  u <- e(s) u + t   or   (u - t) / e(s);   optional compact bf16 copy
//   cpl_mode 2:  y, dy -> x, dx in place; [ds | dt] as bf16; optional bf16 copy of x
// The subnet output never reaches memory.  Rows are 64 contiguous bytes per thread and tensor (16-byte vector accesses).
// Only 8 warps per SM do this math (the standalone kernels spread it over 64), so it uses the fast forms below: a
// degree-7 polynomial in t^2 for atan (1.6e-7 absolute on [0, 1], reciprocal argument beyond 1), ex2.approx, approximate
// division -- errors at fp32 rounding level, far inside the bf16 path's tolerance; the fp32 paths never take this epilogue.
struct CplRegs { float4 u[4], d[4]; };
__device__ __forceinline__ void cpl_prefetch(const CplParams& p, CplRegs& R, long long pix, int ch0, bool row_ok) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const bool ok = row_ok && ch0 + 4 * q < p.L;
    R.u[q] = ok ? *reinterpret_cast<const float4*>(p.u + pix * p.u_stride + ch0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    R.d[q] = (ok && p.mode == 2) ? *reinterpret_cast<const float4*>(p.du + pix * p.du_stride + ch0 + 4 * q)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int TW>
__device__ __forceinline__ void cpl_row(const Params& p, int b, int h0, int w0, int quarter, int lane, bool& row_ok, long long& pix) {
  const int row = quarter * 32 + lane;
  const int oh = h0 + row / TW, ow = w0 + row % TW;
  row_ok = (oh < p.H) && (ow < p.W) && (b < p.B);
  pix = ((long long)b * p.H + oh) * p.W + ow;
}

// One slab: v = 32 accumulator columns (s_k, t_k interleaved) of this thread's pixel, bs = their biases, cur = the 16
// channels of u (and du) loaded earlier; channels ch0 .. ch0+15.
__device__ __forceinline__ void cpl_slab(const CplParams& p, const uint32_t (&v)[32], const float* bs, long long pix, int ch0,
                                         const CplRegs& cur) {
  const int L = p.L;
  const float inv_clamp = 1.0f / p.clamp;
  float* up = p.u + pix * p.u_stride + ch0;
  float* dp = p.du + pix * p.du_stride + ch0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {           // four channels at a time
    if (ch0 + 4 * q < L) {
      const float uu[4] = {cur.u[q].x, cur.u[q].y, cur.u[q].z, cur.u[q].w};
      float y[4];
      if (p.mode == 1) {
        float sk[4], tk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sv = __uint_as_float(v[8 * q + 2 * e]) + bs[8 * q + 2 * e];
          const float tv = __uint_as_float(v[8 * q + 2 * e + 1]) + bs[8 * q + 2 * e + 1];
          float ex, dg;
          glow_scale_fast(p.clamp, inv_clamp, sv, ex, dg);
          y[e] = p.inverse ? __fdividef(uu[e] - tv, ex) : fmaf(ex, uu[e], tv);
          sk[e] = sv; tk[e] = tv;
        }
        *reinterpret_cast<float4*>(up + 4 * q) = make_float4(y[0], y[1], y[2], y[3]);
        if (p.a != nullptr) {
          float* ap = p.a + pix * (2 * L) + ch0 + 4 * q;
          *reinterpret_cast<float4*>(ap) = make_float4(sk[0], sk[1], sk[2], sk[3]);
          *reinterpret_cast<float4*>(ap + L) = make_float4(tk[0], tk[1], tk[2], tk[3]);
        }
      } else {
        const float dy[4] = {cur.d[q].x, cur.d[q].y, cur.d[q].z, cur.d[q].w};
        float dx[4], dsv[4], dtv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sv = __uint_as_float(v[8 * q + 2 * e]) + bs[8 * q + 2 * e];
          const float tv = __uint_as_float(v[8 * q + 2 * e + 1]) + bs[8 * q + 2 * e + 1];
          float ex, dg;
          glow_scale_fast(p.clamp, inv_clamp, sv, ex, dg);
          if (!p.inverse) {                  // y = ex*x + t
            y[e] = __fdividef(uu[e] - tv, ex);
            dx[e] = dy[e] * ex;
            dsv[e] = dy[e] * y[e] * ex * dg;
            dtv[e] = dy[e];
          } else {                           // y = (x - t)/ex
            y[e] = fmaf(uu[e], ex, tv);
            const float qv = __fdividef(dy[e], ex);
            dx[e] = qv;
            dsv[e] = -dy[e] * uu[e] * dg;
            dtv[e] = -qv;
          }
        }
        *reinterpret_cast<float4*>(up + 4 * q) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(dp + 4 * q) = make_float4(dx[0], dx[1], dx[2], dx[3]);
        __nv_bfloat16* da = p.da + pix * (2 * L) + ch0 + 4 * q;
        store4(da, make_float4(dsv[0], dsv[1], dsv[2], dsv[3]));
        store4(da + L, make_float4(dtv[0], dtv[1], dtv[2], dtv[3]));
      }
      if (p.bf16 != nullptr) store4(p.bf16 + pix * L + ch0 + 4 * q, make_float4(y[0], y[1], y[2], y[3]));
    }
  }
}

// `first` holds the operands of this warp's first slab, loaded BEFORE the wait for the accumulator; the operands of the
// warp's next slab are loaded while the current one is computed.  The `nsub` warps of a lane quarter take alternate slabs.
__device__ __forceinline__ void epilogue_tile_coupling(const CplParams& p, const float* bias_s, uint32_t t_base, bool row_ok, long long pix,
                                                       int sub, int nsub, CplRegs& first) {
  const int n_slabs = (2 * p.L + 31) / 32;
  CplRegs cur = first;
  for (int s = sub; s < n_slabs; s += nsub) {
    uint32_t v[32];
    tmem_ld32(t_base + s * 32, v);
    CplRegs nxt;
    if (s + nsub < n_slabs) cpl_prefetch(p, nxt, pix, (s + nsub) * 16, row_ok);
    tmem_ld_wait();
    if (row_ok) cpl_slab(p, v, bias_s + s * 32, pix, s * 16, cur);
    cur = nxt;
  }
}

// Epilogue role of one warp (warps 2..9 of the CTA) of the single-CTA kernels.  TW = tile width in pixels.
template <int TW>
__device__ __forceinline__ void run_epilogue(const Params& p, const CUtensorMap* tmO, Barriers* bars, uint8_t* staging,
                                             float* bias_s, uint32_t tmem_base, int warp, int lane) {
    const int ew = warp - 2;                                   // 0..7
    const int quarter = warp & 3;                              // TMEM lanes 32*quarter .. +31 (hardware rule: warp id % 4)
    const int half = ew >> 2;                                  // which of the two warps of this quarter
    uint8_t* stg = staging + ew * STAGING_BYTES;
    const int etid = threadIdx.x - 64;                         // 0..255 among the epilogue threads
    int acc = 0; uint32_t acc_phase = 0;
    int bias_n0 = -1;
    const bool bias_all = p.Cout <= 256;                       // whole bias vector loaded once: no per-tile barriers
    if (bias_all) epilogue_load_bias(p, bias_s, 0, etid, bias_n0);
    for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      int b, h0, w0, n0;
      tile_coords<TW>(p, t, b, h0, w0, n0);
      if (!bias_all) epilogue_load_bias(p, bias_s, n0, etid, bias_n0);
      mbar_wait(smem_u32(&bars->acc_full[acc]), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ACC_STRIDE;
      epilogue_tile<TW>(p, tmO, stg, bias_all ? bias_s + n0 : bias_s, t_base, b, h0, w0, n0, quarter, half, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_out && lane == 0) bulk_wait_all();               // outstanding tensor stores complete before exit
}

}  // namespace tc
}  // namespace sininn
