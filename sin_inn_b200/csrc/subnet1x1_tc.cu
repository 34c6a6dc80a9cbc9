// Fused 1x1 coupling subnet, forward:   out[p] = W2 * relu(W1 * x[p] + b1) + b2      (per pixel p)
//
// Replaces the two cuDNN launches + ReLU of subnet_conv_1x1 (/root/reference/archs.py:15-17) inside a GLOW
// coupling half (call site archs.py:56-64).  One persistent CTA per SM; a tile is 128 consecutive pixels of the
// channels-last activation matrix [npix][C] (a 1x1 convolution has no spatial structure):
//
//   warp 0     TMA producer: W1 once (stays resident), W2 once (resident) or slab-by-slab through a ring when it
//              does not fit next to everything else, and the x tile of every work item (2..4 stage ring)
//   warp 1     MMA issuer (tcgen05.mma.cta_group::1.kind::f16, M = 128 pixels):
//                 MMA1  D1[128 x hidden] = x  . W1^T     (K = Cin in 16/32/64-channel swizzled slabs)
//                 MMA2  D2[128 x Cout]   = h  . W2^T     (K = hidden, A operand = the bf16 hidden tile in smem)
//              software-pipelined as  MMA1(0); { MMA2(i); MMA1(i+1) }  so the tensor pipe works on the next
//              tile's first layer while the epilogue warps drain this tile's output
//   warps 2-9  epilogue, two warps per TMEM lane quarter:
//                 epi1  D1 -> +b1, ReLU, bf16 -> the hidden tile in shared memory, written directly in the
//                       K-major SWIZZLE_128B layout tcgen05.mma reads (the hidden activation never goes to HBM;
//                       optionally it is ALSO stored with TMA, with its ReLU sign bits, for the backward pass)
//                 epi2  D2 -> +b2 -> fp32 staging (aliases this warp's own rows of the hidden tile, which MMA2 has
//                       finished reading) -> TMA tensor store
// HBM traffic per pixel: Cin*2 B in, Cout*4 B out -- the 2*hidden*2 B hidden round trip of the unfused pair is gone.
#include "tc_epilogue.cuh"

namespace sininn {
namespace tc {

constexpr int S1_MAX_STAGES = 4;
// 16 epilogue warps, four per TMEM lane quarter (ncu: the forward kernel is neither DRAM- (3 %) nor tensor-bound (12 %);
// its tile period is the serial chain MMA1 -> first epilogue (128 x 256 values) -> MMA2 -> second epilogue, and the
// first epilogue took ~3200 cycles on eight warps)
constexpr int S1_EPI_WARPS = 16;
constexpr int S1_SUBS = S1_EPI_WARPS / 4;
constexpr int S1_THREADS = 64 + 32 * S1_EPI_WARPS;

struct S1Barriers {
  uint64_t w1_full;
  uint64_t x_full[S1_MAX_STAGES], x_empty[S1_MAX_STAGES];
  uint64_t w2_full[S1_MAX_STAGES], w2_empty[S1_MAX_STAGES];
  uint64_t d1_full, d2_full, h_full;
  uint32_t tmem_base;
  uint32_t pad;
};

struct S1Params {
  long long npix;
  int num_tiles;
  int cin, hidden, cout;
  int kc1, k1_slabs;            // channels per x / W1 slab (16, 32 or 64) and number of slabs
  int hs;                       // hidden / 64
  int n2pad;                    // Cout padded to 16 (UMMA N of MMA2)
  int x_stages, w2_stages;      // w2_stages == hs  =>  W2 resident
  uint32_t layout1, sbo1;       // UMMA descriptor swizzle code / 8-row stride of the x and W1 slabs
  uint32_t x_slab_bytes, w1_slab_bytes, w2_slab_bytes;
  const float* b1;
  const float* b2;
  uint32_t* bits_out;           // optional ReLU sign bits [npix][hidden/32]
  int store_h;                  // 1: also store the hidden tile (bf16) through tmH
  // data-gradient use of the same pipeline (x := dL/d(out), W1 := conv2 dgrad pack, W2 := conv1 dgrad pack):
  const uint32_t* bits_in;      // non-NULL: first epilogue = zero where the stored ReLU sign bit is 0 (no bias, no ReLU)
  int accumulate;               // 1: out += result (TMA reduce-add) instead of out = result
  CplParams cpl;                // GLOW coupling (apply / backward) in the second epilogue; W2 rows interleaved (pack mode 4)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

__global__ void __launch_bounds__(S1_THREADS, 1)
subnet1x1_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO,
                     const __grid_constant__ CUtensorMap tmH, const S1Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* w1_s = smem_raw + pad;
  uint8_t* w2_s = w1_s + (size_t)p.k1_slabs * p.w1_slab_bytes;
  uint8_t* x_s = w2_s + (size_t)p.w2_stages * p.w2_slab_bytes;
  const uint32_t x_stage_bytes = (uint32_t)p.k1_slabs * p.x_slab_bytes;
  uint8_t* h_s = x_s + (size_t)p.x_stages * x_stage_bytes;                 // [hs][128 rows][128 B]
  S1Barriers* bars = reinterpret_cast<S1Barriers*>(h_s + (size_t)p.hs * 16384);
  float* b1_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [256]
  float* b2_s = b1_s + 256;                                                             // [256]

  // warp index through a shuffle: tells ptxas the role branches below are warp-uniform, which lets it keep the
  // MMA/TMA issue loops on the uniform datapath (without it every tcgen05.mma operand costs an R2UR move)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const bool w2_resident = p.w2_stages == p.hs;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->w1_full), 1);
    for (int s = 0; s < S1_MAX_STAGES; ++s) {
      mbar_init(smem_u32(&bars->x_full[s]), 1);
      mbar_init(smem_u32(&bars->x_empty[s]), 1);
      mbar_init(smem_u32(&bars->w2_full[s]), 1);
      mbar_init(smem_u32(&bars->w2_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->d1_full), 1);
    mbar_init(smem_u32(&bars->d2_full), 1);
    mbar_init(smem_u32(&bars->h_full), S1_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmO);
    if (p.store_h) tma_prefetch_desc(&tmH);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();       // barrier init / TMEM allocation above overlap the previous kernel's tail; global memory from here on
  pdl_trigger();
  if (warp >= 2) {      // biases -> shared memory (zero beyond the real channel counts)
    const int e = threadIdx.x - 64;
    if (e < 256) b1_s[e] = (p.b1 != nullptr && e < p.hidden) ? __ldg(p.b1 + e) : 0.f;
    if (e < 256) {
      const int co = p.cpl.mode != 0 ? ((e & 1) ? p.cpl.L + (e >> 1) : (e >> 1)) : e;      // interleaved rows (s_0, t_0, ...)
      b2_s[e] = (p.b2 != nullptr && e < p.cout) ? __ldg(p.b2 + co) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t d1_tmem = tmem_base, d2_tmem = tmem_base + 256;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      const uint32_t full = smem_u32(&bars->w1_full);
      mbar_expect_tx(full, (uint32_t)p.k1_slabs * p.w1_slab_bytes);
      for (int s = 0; s < p.k1_slabs; ++s) tma_load_2d(smem_u32(w1_s) + s * p.w1_slab_bytes, &tmW1, full, s * p.kc1, 0);
    }
    __syncwarp();
    int xs = 0; uint32_t xph = 0;
    int ws = 0; uint32_t wph = 0;
    bool first = true;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      mbar_wait(smem_u32(&bars->x_empty[xs]), xph ^ 1);
      if (elect_one()) {
        const uint32_t full = smem_u32(&bars->x_full[xs]);
        mbar_expect_tx(full, x_stage_bytes);
        for (int s = 0; s < p.k1_slabs; ++s)
          tma_load_2d(smem_u32(x_s) + xs * x_stage_bytes + s * p.x_slab_bytes, &tmX, full, s * p.kc1, t * 128);
      }
      __syncwarp();
      if (++xs == p.x_stages) { xs = 0; xph ^= 1; }
      if (!w2_resident || first) {
        for (int s = 0; s < p.hs; ++s) {
          mbar_wait(smem_u32(&bars->w2_empty[ws]), wph ^ 1);
          if (elect_one()) {
            const uint32_t full = smem_u32(&bars->w2_full[ws]);
            mbar_expect_tx(full, p.w2_slab_bytes);
            tma_load_2d(smem_u32(w2_s) + ws * p.w2_slab_bytes, &tmW2, full, s * 64, 0);
          }
          __syncwarp();
          if (++ws == p.w2_stages) { ws = 0; wph ^= 1; }
        }
        first = false;
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.hidden >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n2pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t x_desc0 = make_desc(smem_u32(x_s), p.sbo1, p.layout1);
    const uint64_t w1_desc0 = make_desc(smem_u32(w1_s), p.sbo1, p.layout1);
    const uint64_t h_desc0 = make_desc(smem_u32(h_s), 1024, 2);
    const uint64_t w2_desc0 = make_desc(smem_u32(w2_s), 1024, 2);
    const int mma1_per_slab = p.kc1 / 16;
    int xs = 0; uint32_t xph = 0;
    int ws = 0; uint32_t wph = 0;
    uint32_t hph = 0;
    bool first = true;

    auto issue_mma1 = [&]() {
      mbar_wait(smem_u32(&bars->x_full[xs]), xph);
      tc_fence_after();
      if (elect_one()) {
        for (int s = 0; s < p.k1_slabs; ++s) {
          const uint64_t xd = x_desc0 + (uint64_t)((xs * x_stage_bytes + s * p.x_slab_bytes) >> 4);
          const uint64_t wd = w1_desc0 + (uint64_t)((s * p.w1_slab_bytes) >> 4);
          for (int k = 0; k < mma1_per_slab; ++k) umma_bf16(d1_tmem, xd + 2 * k, wd + 2 * k, idesc1, (s | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&bars->x_empty[xs]));
        umma_commit(smem_u32(&bars->d1_full));
      }
      __syncwarp();
      if (++xs == p.x_stages) { xs = 0; xph ^= 1; }
    };

    mbar_wait(smem_u32(&bars->w1_full), 0);
    if ((int)blockIdx.x < p.num_tiles) issue_mma1();
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      mbar_wait(smem_u32(&bars->h_full), hph);       // hidden tile written (and D1 drained) by all epilogue warps
      hph ^= 1;
      tc_fence_after();
      for (int s = 0; s < p.hs; ++s) {
        if (!w2_resident || first) {
          mbar_wait(smem_u32(&bars->w2_full[ws]), wph);
          tc_fence_after();
        }
        const uint64_t hd = h_desc0 + (uint64_t)((s * 16384) >> 4);
        const uint64_t wd = w2_desc0 + (uint64_t)((ws * p.w2_slab_bytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d2_tmem, hd + 2 * k, wd + 2 * k, idesc2, (s | k) != 0 ? 1u : 0u);
          if (!w2_resident) umma_commit(smem_u32(&bars->w2_empty[ws]));
        }
        __syncwarp();
        if (++ws == p.w2_stages) { ws = 0; wph ^= 1; }
      }
      first = false;
      if (elect_one()) umma_commit(smem_u32(&bars->d2_full));
      __syncwarp();
      if (t + (int)gridDim.x < p.num_tiles) issue_mma1();
    }
  } else {
    // ======================= epilogue warps =======================
    const int ew = warp - 2;
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may read (warp id % 4)
    const int half = ew >> 2;                        // which of the S1_SUBS warps of this quarter (0..3)
    const int row = quarter * 32 + lane;             // pixel row inside the tile
    const uint32_t lane_base = ((uint32_t)(quarter * 32) << 16);
    const int bit_words = p.hidden >> 5;
    const int n_out_slabs = (p.cout + 31) / 32;
    int n_pieces = 0;
    for (int s = half; s < p.hs; s += S1_SUBS) ++n_pieces;
    uint32_t dph = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      const long long pix = (long long)t * 128 + row;
      const bool row_ok = pix < p.npix;
      // ---------------- epi1: D1 -> hidden tile (bf16, swizzled K-major) ----------------
      mbar_wait(smem_u32(&bars->d1_full), dph);
      tc_fence_after();
      if (lane == 0) bulk_wait_read0();              // this warp's earlier TMA stores have finished reading its rows
      __syncwarp();
      for (int s = half; s < p.hs; s += S1_SUBS) {
        uint32_t v0[32], v1[32];
        tmem_ld32(d1_tmem + lane_base + s * 64, v0);
        tmem_ld32(d1_tmem + lane_base + s * 64 + 32, v1);
        tmem_ld_wait();
        // packed bf16x2 arithmetic (the epilogue is ALU-pipe bound, tc_epilogue.cuh): pk[j] = elements 2j, 2j+1 of the slab
        uint32_t pk[32];
        if (p.bits_in != nullptr) {                  // gradient through the ReLU: keep where the forward output was > 0
          uint2 mb = make_uint2(0u, 0u);
          if (row_ok) mb = __ldg(reinterpret_cast<const uint2*>(p.bits_in + pix * bit_words + 2 * s));
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            pk[j] = pack_bf16(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1])) & sign_bits_expand(mb.x, j);
            pk[16 + j] = pack_bf16(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])) & sign_bits_expand(mb.y, j);
          }
        } else {
          const float* bs = b1_s + s * 64;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 ba = *reinterpret_cast<const float4*>(bs + 4 * q), bb = *reinterpret_cast<const float4*>(bs + 32 + 4 * q);
            pk[2 * q] = bf16x2_relu(pack_bf16(__uint_as_float(v0[4 * q]) + ba.x, __uint_as_float(v0[4 * q + 1]) + ba.y));
            pk[2 * q + 1] = bf16x2_relu(pack_bf16(__uint_as_float(v0[4 * q + 2]) + ba.z, __uint_as_float(v0[4 * q + 3]) + ba.w));
            pk[16 + 2 * q] = bf16x2_relu(pack_bf16(__uint_as_float(v1[4 * q]) + bb.x, __uint_as_float(v1[4 * q + 1]) + bb.y));
            pk[16 + 2 * q + 1] = bf16x2_relu(pack_bf16(__uint_as_float(v1[4 * q + 2]) + bb.z, __uint_as_float(v1[4 * q + 3]) + bb.w));
          }
        }
        if (p.bits_out != nullptr) {
          uint32_t s0 = 0u, s1 = 0u;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            s0 |= bf16x2_gt0_mask(pk[j]) & ((1u << j) | (1u << (16 + j)));
            s1 |= bf16x2_gt0_mask(pk[16 + j]) & ((1u << j) | (1u << (16 + j)));
          }
          if (row_ok) {
            uint2* dst = reinterpret_cast<uint2*>(p.bits_out + pix * bit_words + 2 * s);
            *dst = make_uint2(s0, s1);
          }
        }
        uint8_t* hrow = h_s + s * 16384 + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(hrow + ((q ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      fence_async_smem();                            // generic-proxy writes -> visible to tcgen05.mma / TMA
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bars->h_full));
        if (p.store_h) {
          for (int s = half; s < p.hs; s += S1_SUBS)
            tma_store_2d(&tmH, smem_u32(h_s) + s * 16384 + quarter * 4096, s * 64, t * 128 + quarter * 32);
          bulk_commit();
        }
      }
      // ---------------- epi2: D2 -> +b2 -> fp32 staging (own rows of the hidden tile) -> TMA store ----------------
      if (p.cpl.mode != 0) {
        // ... or the GLOW half-step on the output held in registers (tc_epilogue.cuh): operands of this warp's first
        // slab are fetched before the wait for MMA2; the subnet output is never stored
        CplRegs cpl;
        cpl_prefetch(p.cpl, cpl, pix, half * 16, row_ok);
        mbar_wait(smem_u32(&bars->d2_full), dph);
        dph ^= 1;
        tc_fence_after();
        epilogue_tile_coupling(p.cpl, b2_s, d2_tmem + lane_base, row_ok, pix, half, S1_SUBS, cpl);
        tc_fence_before();
        continue;
      }
      mbar_wait(smem_u32(&bars->d2_full), dph);
      dph ^= 1;
      tc_fence_after();
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
      int idx = 0;
      // (only warps that own a piece of the hidden tile have rows to stage in: min(S1_SUBS, hidden / 64) per quarter)
      const int n_act = p.hs < S1_SUBS ? p.hs : S1_SUBS;
      for (int o = half; half < n_act && o < n_out_slabs; o += n_act, ++idx) {
        const int piece = half + S1_SUBS * (idx % n_pieces);
        uint8_t* stg = h_s + piece * 16384 + quarter * 4096;
        if (idx > 0) {
          if (lane == 0) { if (n_pieces > 1) bulk_wait_read1(); else bulk_wait_read0(); }
          __syncwarp();
        }
        uint32_t v[32];
        tmem_ld32(d2_tmem + lane_base + o * 32, v);
        tmem_ld_wait();
        const float* bs = b2_s + o * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = *reinterpret_cast<const float4*>(bs + 4 * q);
          *reinterpret_cast<float4*>(stg + lane * 128 + ((q ^ (lane & 7)) << 4)) =
              make_float4(__uint_as_float(v[4 * q]) + bq.x, __uint_as_float(v[4 * q + 1]) + bq.y,
                          __uint_as_float(v[4 * q + 2]) + bq.z, __uint_as_float(v[4 * q + 3]) + bq.w);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.accumulate) tma_reduce_add_2d(&tmO, smem_u32(stg), o * 32, t * 128 + quarter * 32);
          else tma_store_2d(&tmO, smem_u32(stg), o * 32, t * 128 + quarter * 32);
          bulk_commit();
        }
      }
      tc_fence_before();
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static bool encode_2d(EncodeTiledFn encode, CUtensorMap* tm, CUtensorMapDataType dt, int esz, const void* base, long long inner,
                      long long outer, long long stride_elems, int box_inner, int box_outer, CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)stride_elems * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  return encode(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace sininn

using namespace sininn;

// shared-memory plan of the fused kernel: W1 resident, hidden tile, >= 2 W2 slabs, >= 2 x stages
static bool s1_fits(int Cin, int hidden, int n2pad, int k1_pad) {
  const int kc1 = sininn::tc::pick_kc(k1_pad);
  const int k1_slabs = (Cin + kc1 - 1) / kc1;
  const int row1 = kc1 * 2;
  const int budget = 227 * 1024 - 1024 - 256 - 2048;
  const int fixed = k1_slabs * hidden * row1 + (hidden / 64) * 16384;
  const int x_stage = k1_slabs * 128 * row1;
  return fixed + 2 * n2pad * 128 + 2 * x_stage <= budget;
}

extern "C" {

int sininn_subnet1x1_supported(int Cin, int hidden, int Cout) {
  if (hidden % 64 != 0 || hidden < 64 || hidden > 256 || Cout <= 0 || Cout > 256 || Cout % 4 != 0 || Cin <= 0 || Cin % 8 != 0) return 0;
  return s1_fits(Cin, hidden, (Cout + 15) / 16 * 16, (Cin + 15) / 16 * 16) ? 1 : 0;
}

int sininn_subnet1x1_fwd_tc(const sininn_subnet1x1_desc* d, sininn_stream_t stream) {
  using namespace sininn::tc;
  SININN_CHECK_ARG(d != nullptr && d->x && d->w1pack && d->w2pack && (d->out || d->cpl_mode != 0), "subnet1x1: null pointer");
  if (d->cpl_mode != 0) {
    SININN_CHECK_ARG(d->cpl_mode == 1 || d->cpl_mode == 2, "subnet1x1: cpl_mode must be 0, 1 or 2");
    SININN_CHECK_ARG(d->Cout == 2 * d->cpl_L && (d->cpl_L % 8) == 0, "subnet1x1: the coupling epilogue needs Cout = 2 L, L %% 8 == 0 (Cout=%d L=%d)",
                     d->Cout, d->cpl_L);
    SININN_CHECK_ARG(d->cpl_u && aligned16(d->cpl_u) && (d->cpl_u_stride % 4) == 0 && d->cpl_clamp > 0.f, "subnet1x1: bad coupling slice");
    SININN_CHECK_ARG(d->cpl_mode == 1 || (d->cpl_du && aligned16(d->cpl_du) && (d->cpl_du_stride % 4) == 0 && d->cpl_da && aligned8(d->cpl_da)),
                     "subnet1x1: the coupling backward epilogue needs the gradient slice and the [ds | dt] output");
    SININN_CHECK_ARG(!d->cpl_bf16 || aligned8(d->cpl_bf16), "subnet1x1: misaligned bf16 copy");
    SININN_CHECK_ARG(!d->cpl_a || (d->cpl_mode == 1 && aligned16(d->cpl_a)), "subnet1x1: cpl_a needs cpl_mode 1 and 16-byte alignment");
    SININN_CHECK_ARG(!d->mask_bits && !d->accumulate, "subnet1x1: the coupling epilogue is a forward-mode option");
  }
  SININN_CHECK_ARG(d->npix > 0 && d->Cin > 0 && d->Cout > 0, "subnet1x1: bad shape");
  SININN_CHECK_ARG(d->hidden % 64 == 0 && d->hidden >= 64 && d->hidden <= 256, "subnet1x1: hidden width must be 64..256 in steps of 64 (got %d)", d->hidden);
  SININN_CHECK_ARG(d->Cout <= 256 && d->n2_pad % 16 == 0 && d->n2_pad >= d->Cout && d->n2_pad <= 256, "subnet1x1: bad Cout / n2_pad");
  SININN_CHECK_ARG(d->k1_pad % 16 == 0 && d->k1_pad >= d->Cin, "subnet1x1: bad k1_pad");
  SININN_CHECK_ARG(aligned16(d->x) && (d->x_stride * 2) % 16 == 0, "subnet1x1: x must be 16-byte aligned with a pixel stride that is a multiple of 8");
  SININN_CHECK_ARG(d->cpl_mode != 0 || (aligned16(d->out) && (d->out_stride * 4) % 16 == 0),
                   "subnet1x1: out must be 16-byte aligned with a pixel stride that is a multiple of 4");
  SININN_CHECK_ARG(aligned16(d->w1pack) && aligned16(d->w2pack), "subnet1x1: packed weights misaligned");
  SININN_CHECK_ARG(d->h_out == nullptr || (aligned16(d->h_out) && (d->h_stride * 2) % 16 == 0), "subnet1x1: h_out misaligned");
  SININN_CHECK_ARG(d->npix < (1ll << 31) - 256, "subnet1x1: too many pixels");
  SININN_CHECK_ARG(!(d->mask_bits && (d->b1 || d->bits_out)), "subnet1x1: mask_bits (gradient mode) excludes b1 and bits_out");
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("subnet1x1: cuTensorMapEncodeTiled not available from the driver"); return SININN_ECUDA; }

  S1Params p;
  p.npix = d->npix;
  p.num_tiles = (int)((d->npix + 127) / 128);
  p.cin = d->Cin; p.hidden = d->hidden; p.cout = d->Cout;
  p.kc1 = pick_kc(d->k1_pad);
  p.k1_slabs = (d->Cin + p.kc1 - 1) / p.kc1;
  p.hs = d->hidden / 64;
  p.n2pad = d->n2_pad;
  const uint32_t row1 = (uint32_t)p.kc1 * 2;
  p.layout1 = p.kc1 == 64 ? 2u : (p.kc1 == 32 ? 4u : 6u);
  p.sbo1 = 8 * row1;
  p.x_slab_bytes = 128 * row1;
  p.w1_slab_bytes = (uint32_t)d->hidden * row1;
  p.w2_slab_bytes = (uint32_t)p.n2pad * 128u;
  p.b1 = d->b1; p.b2 = d->b2;
  p.cpl.mode = d->cpl_mode; p.cpl.L = d->cpl_L; p.cpl.inverse = d->cpl_inverse; p.cpl.clamp = d->cpl_clamp;
  p.cpl.u = d->cpl_u; p.cpl.u_stride = d->cpl_u_stride; p.cpl.du = d->cpl_du; p.cpl.du_stride = d->cpl_du_stride;
  p.cpl.bf16 = reinterpret_cast<__nv_bfloat16*>(d->cpl_bf16); p.cpl.da = reinterpret_cast<__nv_bfloat16*>(d->cpl_da);
  p.cpl.a = d->cpl_mode == 1 ? d->cpl_a : nullptr;
  p.bits_out = reinterpret_cast<uint32_t*>(d->bits_out);
  p.store_h = d->h_out != nullptr ? 1 : 0;
  p.bits_in = reinterpret_cast<const uint32_t*>(d->mask_bits);
  p.accumulate = d->accumulate;
  // shared-memory plan: W1 resident, hidden tile, then as much of W2 as fits (all of it => resident), x ring
  const int budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - 2048 /*biases*/;
  const int fixed = p.k1_slabs * (int)p.w1_slab_bytes + p.hs * 16384;
  const int x_stage = p.k1_slabs * (int)p.x_slab_bytes;
  int w2_stages = p.hs, x_stages = 2;
  if (fixed + w2_stages * (int)p.w2_slab_bytes + x_stages * x_stage > budget) {
    w2_stages = (budget - fixed - x_stages * x_stage) / (int)p.w2_slab_bytes;
    if (w2_stages >= p.hs) w2_stages = p.hs - 1;
    if (w2_stages > S1_MAX_STAGES) w2_stages = S1_MAX_STAGES;
    SININN_CHECK_ARG(w2_stages >= 2, "subnet1x1: does not fit in shared memory (Cin=%d hidden=%d Cout=%d)", d->Cin, d->hidden, d->Cout);
  } else {
    while (x_stages < S1_MAX_STAGES && fixed + w2_stages * (int)p.w2_slab_bytes + (x_stages + 1) * x_stage <= budget) ++x_stages;
  }
  p.w2_stages = w2_stages; p.x_stages = x_stages;

  const CUtensorMapSwizzle swz1 = p.kc1 == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc1 == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap tmX, tmW1, tmW2, tmO, tmH;
  if (!encode_2d(encode, &tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->x, d->Cin, d->npix, d->x_stride, p.kc1, 128, swz1)) {
    set_error("subnet1x1: tensor map (x) failed (Cin=%d stride=%d npix=%lld)", d->Cin, d->x_stride, d->npix); return SININN_ECUDA;
  }
  if (!encode_2d(encode, &tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->w1pack, d->k1_pad, d->hidden, d->k1_pad, p.kc1, d->hidden, swz1)) {
    set_error("subnet1x1: tensor map (W1) failed"); return SININN_ECUDA;
  }
  if (!encode_2d(encode, &tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->w2pack, d->hidden, p.n2pad, d->hidden, 64, p.n2pad, CU_TENSOR_MAP_SWIZZLE_128B)) {
    set_error("subnet1x1: tensor map (W2) failed"); return SININN_ECUDA;
  }
  if (d->cpl_mode != 0) {
    tmO = tmX;                 // (never used: the coupling epilogue consumes the output in registers)
  } else if (!encode_2d(encode, &tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->out, d->Cout, d->npix, d->out_stride, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) {
    set_error("subnet1x1: tensor map (out) failed (Cout=%d stride=%d)", d->Cout, d->out_stride); return SININN_ECUDA;
  }
  if (p.store_h) {
    if (!encode_2d(encode, &tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->h_out, d->hidden, d->npix, d->h_stride, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)) {
      set_error("subnet1x1: tensor map (h_out) failed"); return SININN_ECUDA;
    }
  } else {
    tmH = tmO;
  }
  const size_t smem = (size_t)fixed + (size_t)w2_stages * p.w2_slab_bytes + (size_t)x_stages * x_stage + 256 + 2048 + 1024;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(subnet1x1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("subnet1x1: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SININN_ECUDA; }
    attr_set[dev] = true;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  launch_k(subnet1x1_fwd_kernel, dim3((unsigned)grid), dim3(S1_THREADS), smem, as_stream(stream), tmX, tmW1, tmW2, tmO, tmH, p);
  SININN_CHECK_LAUNCH("subnet1x1_fwd");
  return SININN_OK;
}

}  // extern "C"
