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
#include <stdlib.h>
#include "tc_epilogue.cuh"

namespace sininn {
namespace tc {

int launch_conv_halo(const sininn_conv_desc* d, Params p, int base_offset_mode, cudaStream_t st);   // conv_tc_halo.cu
int launch_conv_pair(const sininn_conv_desc* d, Params p, cudaStream_t st);                          // conv_tc_pair.cu

constexpr int TILE_H = 8, TILE_W = 16, TILE_M = TILE_H * TILE_W;   // 128 pixels = UMMA M

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

  // warp index through a shuffle: tells ptxas the role branches below are warp-uniform, which lets it keep the
  // MMA/TMA issue loops on the uniform datapath (without it every tcgen05.mma operand costs an R2UR move)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
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
  pdl_wait();       // prologue above overlaps the previous kernel's tail; no global memory is touched before this
  pdl_trigger();
  const int k_steps = p.taps * p.k_chunks;

  if (warp == 0) {
    // ======================= TMA producer (whole warp loops, one elected lane issues) =======================
    {
      int stage = 0; uint32_t phase = 0;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        int b, h0, w0, n0;
        tile_coords<TILE_W>(p, t, b, h0, w0, n0);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
            if (elect_one()) {
              const uint32_t full = smem_u32(&bars->full[stage]);
              mbar_expect_tx(full, p.tx_bytes);
              const uint32_t a_dst = ring_u32 + stage * stage_bytes;
              tma_load_4d(a_dst, &tmA, full, kc * p.kc, w0 + dx, h0 + dy, b);
              tma_load_3d(a_dst + p.a_bytes, &tmB, full, kc * p.kc, n0, tap);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (whole warp loops, one elected lane issues) =======================
    {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const int mma_per_step = p.kc / 16;
      // descriptors are built once; inside the loops only the 14-bit start-address field (>>4) is advanced, so the
      // single issuing thread spends a handful of instructions per MMA instead of re-encoding two descriptors
      const uint64_t a_desc0 = make_desc(ring_u32, p.sbo, p.layout_type);
      const uint64_t b_desc0 = make_desc(ring_u32 + p.a_bytes, p.sbo, p.layout_type);
      const uint32_t stage_step = stage_bytes >> 4;
      for (long long t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(smem_u32(&bars->acc_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int ks = 0; ks < k_steps; ++ks) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          tc_fence_after();
          const uint64_t adesc = a_desc0 + (uint64_t)(stage * stage_step);
          const uint64_t bdesc = b_desc0 + (uint64_t)(stage * stage_step);
          if (elect_one()) {
            umma_bf16(d_tmem, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
            for (int k = 1; k < mma_per_step; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
            umma_commit(smem_u32(&bars->empty[stage]));      // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(smem_u32(&bars->acc_full[acc]));   // accumulator complete
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    run_epilogue<TILE_W>(p, &tmO, bars, staging, bias_s, tmem_base, warp, lane);
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
  SININN_CHECK_ARG(d != nullptr && d->in && d->wpack && (d->out || d->cpl_mode != 0), "conv_tc: null pointer");
  if (d->cpl_mode != 0) {
    SININN_CHECK_ARG(d->cpl_mode == 1 || d->cpl_mode == 2, "conv_tc: cpl_mode must be 0, 1 or 2");
    SININN_CHECK_ARG(d->taps == 9 && d->Cout == 2 * d->cpl_L && d->rows_pad <= 256 && (d->cpl_L % 8) == 0,
                     "conv_tc: the fused coupling epilogue needs a 3x3 convolution with Cout = 2 L <= 256, L %% 8 == 0 (Cout=%d L=%d)",
                     d->Cout, d->cpl_L);
    SININN_CHECK_ARG(d->cpl_u && aligned16(d->cpl_u) && (d->cpl_u_stride % 4) == 0 && d->cpl_clamp > 0.f, "conv_tc: bad coupling slice");
    SININN_CHECK_ARG(d->cpl_mode == 1 || (d->cpl_du && aligned16(d->cpl_du) && (d->cpl_du_stride % 4) == 0 && d->cpl_da && aligned8(d->cpl_da)),
                     "conv_tc: the coupling backward epilogue needs the gradient slice and the [ds | dt] output");
    SININN_CHECK_ARG(!d->cpl_bf16 || aligned8(d->cpl_bf16), "conv_tc: misaligned bf16 copy");
    SININN_CHECK_ARG(!d->cpl_a || (d->cpl_mode == 1 && aligned16(d->cpl_a)), "conv_tc: cpl_a needs cpl_mode 1 and 16-byte alignment");
    SININN_CHECK_ARG(d->act == SININN_ACT_NONE && !d->mask && !d->mask_bits && !d->bits_out && !d->accumulate && d->alpha == 1.0f,
                     "conv_tc: the coupling epilogue replaces every other epilogue option");
  }
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
  p.cpl.mode = d->cpl_mode; p.cpl.L = d->cpl_L; p.cpl.inverse = d->cpl_inverse; p.cpl.clamp = d->cpl_clamp;
  p.cpl.u = d->cpl_u; p.cpl.u_stride = d->cpl_u_stride; p.cpl.du = d->cpl_du; p.cpl.du_stride = d->cpl_du_stride;
  p.cpl.bf16 = reinterpret_cast<__nv_bfloat16*>(d->cpl_bf16); p.cpl.da = reinterpret_cast<__nv_bfloat16*>(d->cpl_da);
  p.cpl.a = d->cpl_mode == 1 ? d->cpl_a : nullptr;
  const int esz = p.out_f32 ? 4 : 2;
  // TMA epilogue needs a 16-byte aligned output slice / pixel stride and no per-element mask tensor
  p.tma_out = (d->out && aligned16(d->out) && ((long long)d->out_stride * esz) % 16 == 0 && d->mask == nullptr) ? 1 : 0;

  // 3x3 with 64-channel slabs: halo-reuse kernel (activation patch loaded once per slab instead of nine times)
  // (SININN_HALO=0 disables it; measured on B200: the UMMA swizzle XOR is taken from the absolute shared-memory
  //  address bits, so a descriptor may start at any 128-byte row of the TMA-written halo box with base_offset 0)
  // 3x3: CTA-pair kernel (cta_group::2, weight rows split over the pair, resident when they fit); SININN_PAIR=0 disables
  static int pair_mode = -1;
  if (pair_mode < 0) {
    const char* e = getenv("SININN_PAIR");
    pair_mode = e ? atoi(e) : 1;
  }
  if (pair_mode > 0 && d->taps == 9) {
    const int rc = launch_conv_pair(d, p, as_stream(stream));
    if (rc != SININN_EUNSUPPORTED) return rc;
  }
  if (d->cpl_mode != 0) {
    set_error("conv_tc: the fused coupling epilogue is a CTA-pair kernel feature (shape not taken, or SININN_PAIR=0)");
    return SININN_EUNSUPPORTED;
  }
  static int halo_mode = -1;
  if (halo_mode < 0) {
    const char* e = getenv("SININN_HALO");
    halo_mode = e ? atoi(e) : 1;
  }
  if (halo_mode > 0 && d->taps == 9 && p.kc == 64)
    return launch_conv_halo(d, p, 0, as_stream(stream));

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
  launch_k(conv_tc_kernel, dim3((unsigned)grid), dim3(NUM_THREADS), smem, as_stream(stream), tmA, tmB, tmO, p);
  SININN_CHECK_LAUNCH("conv_tc");
  return SININN_OK;
}

}  // extern "C"
