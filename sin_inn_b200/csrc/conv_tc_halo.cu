// 3x3 implicit-GEMM convolution with HALO REUSE (tcgen05 / TMEM / TMA, sm_100a).
//
// conv_tc.cu loads the activation patch nine times per 64-channel slab (once per filter tap, shifted by TMA).
// Here the patch is loaded ONCE per slab with its one-pixel halo -- box {64 ch, 16 w, 18 h}: a 16x8-pixel tile plus
// halo, 16 columns wide so that consecutive tile rows are exactly 16 x 128 B = 2048 B apart in shared memory --
// and every tap is just a different UMMA shared-memory descriptor into the same buffer:
//     start = base + ((1+dy)*16 + (1+dx)) * 128 B,   SBO (8-row group stride) = 2048 B,
// with base_offset 0: on B200 the UMMA 128B-swizzle XOR is computed from the absolute shared-memory address bits
// (verified against the CPU oracle; setting base_offset = (start >> 7) & 7 double-counts the phase and is wrong),
// which is also how TMA laid the box out.  L2 -> SM activation traffic drops from 9x to 2.25x of the tensor; the weight slabs still stream per
// (slab, tap) through their own ring.  Tile = 16 rows x 8 columns (UMMA M = 128).
#include "tc_epilogue.cuh"

namespace sininn {
namespace tc {

constexpr int HTILE_W = 8, HTILE_H = 16;
constexpr int HALO_W = 16, HALO_H = HTILE_H + 2;
constexpr uint32_t HALO_BYTES = HALO_W * HALO_H * 128;          // 36864 (64 channels x 2 B per halo pixel)
constexpr int MAX_A_STAGES = 4;

struct HaloBarriers {
  uint64_t a_full[MAX_A_STAGES];
  uint64_t a_empty[MAX_A_STAGES];
};

struct HaloParams {
  Params p;
  int a_stages, b_stages;
  int base_offset_mode;       // 1: descriptor base_offset = (start >> 7) & 7; 0: leave it 0
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const HaloParams hp) {
  const Params& p = hp.p;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* a_ring = smem_raw + pad;                                          // [a_stages][HALO_BYTES]
  uint8_t* b_ring = a_ring + (size_t)hp.a_stages * HALO_BYTES;               // [b_stages][b_bytes]
  uint8_t* staging = b_ring + (size_t)hp.b_stages * p.b_bytes;               // 1024-aligned
  Barriers* bars = reinterpret_cast<Barriers*>(staging + NUM_EPI_WARPS * STAGING_BYTES);
  float* bias_s = reinterpret_cast<float*>(staging + NUM_EPI_WARPS * STAGING_BYTES + BARRIER_BYTES);
  HaloBarriers* hb = reinterpret_cast<HaloBarriers*>(staging + NUM_EPI_WARPS * STAGING_BYTES + BARRIER_BYTES + 1024);

  // warp index through a shuffle: tells ptxas the role branches below are warp-uniform, which lets it keep the
  // MMA/TMA issue loops on the uniform datapath (without it every tcgen05.mma operand costs an R2UR move)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t a_u32 = smem_u32(a_ring), b_u32 = smem_u32(b_ring);

  if (threadIdx.x == 0) {
    for (int s = 0; s < hp.b_stages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int s = 0; s < hp.a_stages; ++s) {
      mbar_init(smem_u32(&hb->a_full[s]), 1);
      mbar_init(smem_u32(&hb->a_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bars->acc_full[a]), 1);
      mbar_init(smem_u32(&bars->acc_empty[a]), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_out) tma_prefetch_desc(&tmO);
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
    // ======================= TMA producer (whole warp loops, one elected lane issues) =======================
    {
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pb = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int b, h0, w0, n0;
        tile_coords<HTILE_W>(p, t, b, h0, w0, n0);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(smem_u32(&hb->a_empty[sa]), pa ^ 1);
          if (elect_one()) {
            const uint32_t afull = smem_u32(&hb->a_full[sa]);
            mbar_expect_tx(afull, HALO_BYTES);
            tma_load_4d(a_u32 + sa * HALO_BYTES, &tmA, afull, kc * 64, w0 - 1, h0 - 1, b);
          }
          __syncwarp();
          if (++sa == hp.a_stages) { sa = 0; pa ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(smem_u32(&bars->empty[sb]), pb ^ 1);
            if (elect_one()) {
              const uint32_t bfull = smem_u32(&bars->full[sb]);
              mbar_expect_tx(bfull, (uint32_t)p.n_tile * 128u);
              tma_load_3d(b_u32 + sb * p.b_bytes, &tmB, bfull, kc * 64, n0, tap);
            }
            __syncwarp();
            if (++sb == hp.b_stages) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp loops, one elected lane issues) =======================
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pb = 0;
      int acc = 0; uint32_t acc_phase = 0;
      // descriptors built once; per MMA only the start-address field (units of 16 B) is advanced
      const uint64_t a_desc0 = make_desc(a_u32, HALO_W * 128u, 2);
      const uint64_t b_desc0 = make_desc(b_u32, 1024, 2);
      const uint32_t a_step = HALO_BYTES >> 4, b_step = p.b_bytes >> 4;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(smem_u32(&bars->acc_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(smem_u32(&hb->a_full[sa]), pa);
          tc_fence_after();
          const uint64_t a_stage = a_desc0 + (uint64_t)(sa * a_step);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(smem_u32(&bars->full[sb]), pb);
            tc_fence_after();
            // tap (dy, dx) in 0..2 (halo origin is (h0-1, w0-1)): start row (dy*16 + dx) of the halo box, 8 x 16 B per row
            const uint64_t adesc = a_stage + (uint64_t)(((tap / 3) * HALO_W + (tap % 3)) * 8);
            const uint64_t bdesc = b_desc0 + (uint64_t)(sb * b_step);
            if (elect_one()) {
              umma_bf16(d_tmem, adesc, bdesc, idesc, (kc | tap) != 0 ? 1u : 0u);
              umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
              umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
              umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
              umma_commit(smem_u32(&bars->empty[sb]));
            }
            __syncwarp();
            if (++sb == hp.b_stages) { sb = 0; pb ^= 1; }
          }
          if (elect_one()) umma_commit(smem_u32(&hb->a_empty[sa]));   // halo buffer free once all nine taps retired
          __syncwarp();
          if (++sa == hp.a_stages) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) umma_commit(smem_u32(&bars->acc_full[acc]));
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    run_epilogue<HTILE_W>(p, &tmO, bars, staging, bias_s, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Host launcher, called by sininn_conv_tc when the shape qualifies (3x3, Cin slabs of 64).
int launch_conv_halo(const sininn_conv_desc* d, Params p, int base_offset_mode, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    set_error("conv_tc(halo): cuTensorMapEncodeTiled not available from the driver");
    return SININN_ECUDA;
  }
  p.tiles_h = (d->H + HTILE_H - 1) / HTILE_H;
  p.tiles_w = (d->W + HTILE_W - 1) / HTILE_W;
  p.num_tiles = (long long)d->B * p.tiles_h * p.tiles_w * p.n_tiles;
  HaloParams hp;
  hp.base_offset_mode = base_offset_mode;
  const int budget = SMEM_RING_BUDGET - 1024 - (int)sizeof(HaloBarriers);
  int a_stages = 3, b_stages;
  for (;; --a_stages) {
    b_stages = (budget - a_stages * (int)HALO_BYTES) / (int)p.b_bytes;
    if (b_stages >= 3 || a_stages == 2) break;
  }
  if (b_stages > MAX_STAGES) b_stages = MAX_STAGES;
  if (b_stages < 2) {
    set_error("conv_tc(halo): tile does not fit in shared memory");
    return SININN_EINVAL;
  }
  hp.a_stages = a_stages; hp.b_stages = b_stages;
  hp.p = p;
  const int esz = p.out_f32 ? 4 : 2;
  CUtensorMap tmA, tmB, tmO;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_stride * 2, (cuuint64_t)d->W * d->in_stride * 2,
                             (cuuint64_t)d->H * d->W * d->in_stride * 2};
    cuuint32_t box[4] = {64, HALO_W, HALO_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->in), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc(halo): tensor map (activations) failed with %d", (int)r); return SININN_ECUDA; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d->k_pad, (cuuint64_t)d->rows_pad, (cuuint64_t)d->taps};
    cuuint64_t strides[2] = {(cuuint64_t)d->k_pad * 2, (cuuint64_t)d->rows_pad * d->k_pad * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)p.n_tile, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->wpack), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc(halo): tensor map (weights) failed with %d", (int)r); return SININN_ECUDA; }
  }
  if (p.tma_out) {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t strides[3] = {(cuuint64_t)d->out_stride * esz, (cuuint64_t)d->W * d->out_stride * esz,
                             (cuuint64_t)d->H * d->W * d->out_stride * esz};
    cuuint32_t box[4] = {(cuuint32_t)(128 / esz), HTILE_W, 32 / HTILE_W, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmO, p.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d->out, dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc(halo): tensor map (output) failed with %d", (int)r); return SININN_ECUDA; }
  } else {
    tmO = tmA;
  }
  const size_t smem = (size_t)a_stages * HALO_BYTES + (size_t)b_stages * p.b_bytes + EPI_SMEM_BYTES + BARRIER_BYTES + 1024 +
                      sizeof(HaloBarriers) + 1024;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_tc(halo): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SININN_ECUDA; }
    attr_set[dev] = true;
  }
  long long grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  launch_k(conv_tc_halo_kernel, dim3((unsigned)grid), dim3(NUM_THREADS), smem, st, tmA, tmB, tmO, hp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("conv_tc(halo): launch failed: %s", cudaGetErrorString(e)); return SININN_ECUDA; }
  return SININN_OK;
}

}  // namespace tc
}  // namespace sininn
