// CUDA-core implicit-GEMM convolution (fprop / dgrad), weight gradient and weight packing.
// This is the fp32-accurate path (parity 1e-4 against the reference) and the on-device
// cross-check for the tcgen05 path in conv_tc.cu; it accepts fp32 or bf16 operands and
// always accumulates in fp32.
//
// Replaces nn.Conv2d in subnet_conv / subnet_conv_1x1 (/root/reference/archs.py:11-17) and
// DenseBlock (/root/reference/archs.py:77-81,88-95) and their autograd backward.
#include "common.cuh"

namespace sininn {

__device__ __forceinline__ bool aligned_dev16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int BM = 64;   // pixels per CTA tile (8 x 8 spatial patch)
constexpr int BN = 64;   // output channels per CTA tile
constexpr int BK = 16;   // reduction slice
constexpr int TPX = 8;   // tile is TPX x TPX pixels

struct ConvArgs {
  int B, H, W, Cin, Cout, taps;
  const void* in; int in_stride;
  const void* w; int rows_pad, k_pad;
  const float* bias;
  void* out; int out_stride;
  int act; float slope;
  const void* mask; int mask_stride; int mask_act;
  int accumulate; float alpha;
  int tiles_h, tiles_w;
  int vec_in;    // 1: input rows may be read 4 channels at a time
};

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const TI* __restrict__ in = reinterpret_cast<const TI*>(a.in);
  const TI* __restrict__ wp = reinterpret_cast<const TI*>(a.w);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  int tile = blockIdx.x;
  const int tw = tile % a.tiles_w; tile /= a.tiles_w;
  const int th = tile % a.tiles_h;
  const int b = tile / a.tiles_h;
  const int h0 = th * TPX, w0 = tw * TPX;
  const int n0 = blockIdx.y * BN;

  // loader roles: A: pixel lm, 4 channels at lk; B: out-channel ln, 4 k at lk
  const int lm = tid >> 2, lk = (tid & 3) * 4;
  const int lh = h0 + (lm >> 3), lw = w0 + (lm & 7);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < a.taps; ++tap) {
    const int dy = (a.taps == 9) ? tap / 3 - 1 : 0;
    const int dx = (a.taps == 9) ? tap % 3 - 1 : 0;
    const int ih = lh + dy, iw = lw + dx;
    const bool pix_ok = (lh < a.H) && (lw < a.W) && ih >= 0 && ih < a.H && iw >= 0 && iw < a.W;
    const TI* src = in + (((long long)b * a.H + ih) * a.W + iw) * (long long)a.in_stride;
    const TI* wrow = wp + ((long long)tap * a.rows_pad + (n0 + lm)) * (long long)a.k_pad;
    const bool row_ok = (n0 + lm) < a.rows_pad;
    for (int c0 = 0; c0 < a.Cin; c0 += BK) {
      // ---- A tile: As[k][m] = in[pixel m shifted by tap][c0+k]
      float av[4] = {0.f, 0.f, 0.f, 0.f};
      if (pix_ok) {
        const int c = c0 + lk;
        if (a.vec_in && c + 3 < a.Cin) {
          float4 t = load4(src + c);
          av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (c + e < a.Cin) av[e] = to_f32(src[c + e]);
        }
      }
      // ---- B tile: Bs[k][n] = w[tap][n0+n][c0+k]   (rows zero-padded to k_pad, k_pad % 4 == 0)
      float bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (row_ok) {
        const int c = c0 + lk;
        if (c + 3 < a.k_pad) {
          float4 t = load4(wrow + c);
          bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (c + e < a.k_pad) bv[e] = to_f32(wrow[c + e]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { As[lk + e][lm] = av[e]; Bs[lk + e][lm] = bv[e]; }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 x = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 y = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float xa[4] = {x.x, x.y, x.z, x.w}, yb[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], yb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  TO* __restrict__ out = reinterpret_cast<TO*>(a.out);
  const TO* __restrict__ mask = reinterpret_cast<const TO*>(a.mask);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = ty * 4 + i;
    const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
    if (oh >= a.H || ow >= a.W) continue;
    const long long pix = ((long long)b * a.H + oh) * a.W + ow;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= a.Cout) continue;
      float v = acc[i][j] + (a.bias ? a.bias[co] : 0.f);
      v = act_fwd(a.act, a.slope, v);
      if (mask) v *= act_grad(a.mask_act, a.slope, to_f32(mask[pix * a.mask_stride + co]));
      v *= a.alpha;
      TO* o = out + pix * a.out_stride + co;
      if (a.accumulate) v += to_f32(*o);
      *o = from_f32<TO>(v);
    }
  }
}


// ---------------------------------------------------------------- fp32-accurate tensor-core path ("split" operands)
// An fp32 value x is carried as three bf16 terms h = bf16(x), m = bf16(x - h), l = bf16(x - h - m) (24 mantissa bits),
// a weight w likewise as wh, wm, wl.  Activations are laid out as SIX channel blocks [h | m | h | l | m | h] and weights
// as [wh | wh | wm | wh | wm | wl], so that the ordinary bf16 implicit GEMM over the 6x longer K computes every product
// term down to second order,
//     h*wh + m*wh + h*wm + l*wh + m*wm + h*wl  =  x*w - (third-order terms ~2^-26 |x w|),
// i.e. fp32-level accuracy with fp32 accumulation in TMEM.  (Four blocks with two-term weights are 2^-17 accurate,
// which passes value tolerances but flips ReLU decisions of near-zero hidden units ~30x more often than fp32 does,
// and a flipped unit is a finite jump in the gradients -- measured on the parity nets, see DESIGN.md.)
constexpr int SPLIT_BLOCKS = 6;
__device__ __forceinline__ float split_weight_term(float w, int block) {
  const float wh = __bfloat162float(__float2bfloat16_rn(w));
  if (block == 0 || block == 1 || block == 3) return wh;
  const float r1 = w - wh;
  const float wm = __bfloat162float(__float2bfloat16_rn(r1));
  return block == 5 ? r1 - wm : wm;       // the caller rounds to bf16
}

// ---------------------------------------------------------------- weight packing
template <typename TO>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int mode,
                                                          TO* __restrict__ out, int rows_pad, int k_pad) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)taps * rows_pad * k_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int k = (int)(idx % k_pad);
    long long r = idx / k_pad;
    int row = (int)(r % rows_pad);
    int tap = (int)(r / rows_pad);
    float v = 0.f;
    int blockk = 0;
    if (mode == 2 || mode == 3) { const int kp = k_pad / SPLIT_BLOCKS; blockk = k / kp; k = k % kp; }   // split packs: six K blocks
    if (mode == 4) {            // fprop, output rows interleaved: (s_0, t_0, s_1, t_1, ...)
      const int co = (row & 1) ? Cout / 2 + (row >> 1) : (row >> 1);
      if (row < Cout && k < Cin) v = w[((long long)co * Cin + k) * taps + tap];
    } else if ((mode & 1) == 0) {      // fprop: row = co, k = ci
      if (row < Cout && k < Cin) v = w[((long long)row * Cin + k) * taps + tap];
    } else {                    // dgrad: row = ci, k = co, spatially flipped
      if (row < Cin && k < Cout) v = w[((long long)k * Cin + row) * taps + (taps - 1 - tap)];
    }
    if (mode == 2 || mode == 3) v = split_weight_term(v, blockk);
    out[idx] = from_f32<TO>(v);
  }
}

// all conv weights of a network in ONE launch: jobs[j] = {src, dst, Cout, Cin, taps, mode, rows_pad, k_pad}
template <typename TO>
__global__ void __launch_bounds__(256) pack_weight_batched_kernel(const long long* __restrict__ jobs) {
  pdl_wait();
  pdl_trigger();
  const long long* j = jobs + (long long)blockIdx.y * 8;
  const float* w = reinterpret_cast<const float*>(j[0]);
  TO* out = reinterpret_cast<TO*>(j[1]);
  const int Cout = (int)j[2], Cin = (int)j[3], taps = (int)j[4], mode = (int)j[5], rows_pad = (int)j[6], k_pad = (int)j[7];
  const long long total = (long long)taps * rows_pad * k_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int k = (int)(idx % k_pad);
    long long r = idx / k_pad;
    int row = (int)(r % rows_pad);
    int tap = (int)(r / rows_pad);
    float v = 0.f;
    int blockk = 0;
    if (mode == 2 || mode == 3) { const int kp = k_pad / SPLIT_BLOCKS; blockk = k / kp; k = k % kp; }
    if (mode == 4) {
      const int co = (row & 1) ? Cout / 2 + (row >> 1) : (row >> 1);
      if (row < Cout && k < Cin) v = w[((long long)co * Cin + k) * taps + tap];
    } else if ((mode & 1) == 0) {
      if (row < Cout && k < Cin) v = w[((long long)row * Cin + k) * taps + tap];
    } else {
      if (row < Cin && k < Cout) v = w[((long long)k * Cin + row) * taps + (taps - 1 - tap)];
    }
    if (mode == 2 || mode == 3) v = split_weight_term(v, blockk);
    out[idx] = from_f32<TO>(v);
  }
}

// ---------------------------------------------------------------- weight gradient
// partial[split][tap][co][ci] = sum over the split's pixels of dy[p][co] * x[p+off(tap)][ci]
struct WgradArgs {
  int B, H, W, Cin, Cout, taps;
  const void* x; int x_stride;
  const void* dy; int dy_stride;
  float* partial;
  long long npix, pix_per_split;
  int ci_tiles;
  int vec_x, vec_dy;
};

template <typename TX, typename TD>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradArgs a) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[BK][BM + 4];   // dy  [k = pixel][co]
  __shared__ float Bs[BK][BN + 4];   // x   [k = pixel][ci]
  const TX* __restrict__ x = reinterpret_cast<const TX*>(a.x);
  const TD* __restrict__ dy = reinterpret_cast<const TD*>(a.dy);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int co0 = blockIdx.x * BM;
  const int tap = blockIdx.y / a.ci_tiles;
  const int ci0 = (blockIdx.y % a.ci_tiles) * BN;
  const int split = blockIdx.z;
  const int oy = (a.taps == 9) ? tap / 3 - 1 : 0;
  const int ox = (a.taps == 9) ? tap % 3 - 1 : 0;
  const long long p0 = split * a.pix_per_split;
  long long p1 = p0 + a.pix_per_split;
  if (p1 > a.npix) p1 = a.npix;

  const int lk = tid >> 4, lc = (tid & 15) * 4;   // loader: pixel lk of the slice, 4 channels at lc
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pb = p0; pb < p1; pb += BK) {
    const long long p = pb + lk;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (p < p1) {
      const int w_ = (int)(p % a.W);
      const long long r = p / a.W;
      const int h_ = (int)(r % a.H);
      const int ih = h_ + oy, iw = w_ + ox;
      if (ih >= 0 && ih < a.H && iw >= 0 && iw < a.W) {
        const TD* dsrc = dy + p * a.dy_stride;
        const int co = co0 + lc;
        if (a.vec_dy && co + 3 < a.Cout) {
          float4 t = load4(dsrc + co);
          av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (co + e < a.Cout) av[e] = to_f32(dsrc[co + e]);
        }
        const TX* xsrc = x + (p + (long long)oy * a.W + ox) * a.x_stride;
        const int ci = ci0 + lc;
        if (a.vec_x && ci + 3 < a.Cin) {
          float4 t = load4(xsrc + ci);
          bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (ci + e < a.Cin) bv[e] = to_f32(xsrc[ci + e]);
        }
      }
    }
    *reinterpret_cast<float4*>(&As[lk][lc]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[lk][lc]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 u = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 v = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float ua[4] = {u.x, u.y, u.z, u.w}, vb[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ua[i], vb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* part = a.partial + ((long long)split * a.taps + tap) * a.Cout * (long long)a.Cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= a.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < a.Cin) part[(long long)co * a.Cin + ci] = acc[i][j];
    }
  }
}

// dw[co][ci][tap] (+)= sum_split partial[split][tap][co][ci]   (fixed order => deterministic)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int Cout, int Cin,
                                                           float* __restrict__ dw, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long per = (long long)taps * Cout * Cin;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < per;
       idx += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(idx % Cin);
    long long r = idx / Cin;
    int co = (int)(r % Cout);
    int tap = (int)(r / Cout);
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[k * per + idx];
    float* o = dw + ((long long)co * Cin + ci) * taps + tap;
    *o = accumulate ? *o + s : s;
  }
}

static inline int wgrad_splits(const sininn_wgrad_desc* d) {
  long long tiles = (long long)((d->Cout + BM - 1) / BM) * ((d->Cin + BN - 1) / BN) * d->taps;
  long long want = ((long long)sm_count() * 8 + tiles - 1) / tiles;
  long long npix = (long long)d->B * d->H * d->W;
  long long max_by_pix = (npix + 255) / 256;
  if (want > max_by_pix) want = max_by_pix;
  if (want > 64) want = 64;
  if (want < 1) want = 1;
  return (int)want;
}

// out[p][j * Lp + c] = term_j(scale * in[p][c]) for the six blocks [h | m | h | l | m | h]; columns c >= L of a block are zero
__global__ void __launch_bounds__(256) split_slice_kernel(const float* __restrict__ in, int in_stride, long long npix, int L, int Lp,
                                                          float scale, __nv_bfloat16* __restrict__ out, int blocks) {
  pdl_wait();
  pdl_trigger();
  const int Lv = Lp / 4;
  const long long total = npix * Lv;
  const bool vec = (in_stride % 4) == 0 && aligned_dev16(in);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Lv) * 4;
    const long long p = idx / Lv;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = in + p * in_stride + c;
    if (vec && c + 3 < L) {
      const float4 t = *reinterpret_cast<const float4*>(src);
      x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (c + e < L) x[e] = src[e];
    }
    float h[4], m[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = x[e] * scale;
      h[e] = __bfloat162float(__float2bfloat16_rn(v));
      const float r1 = v - h[e];
      m[e] = __bfloat162float(__float2bfloat16_rn(r1));
      l[e] = r1 - m[e];
    }
    __nv_bfloat16* o = out + p * (long long)(blocks * Lp) + c;
    store4(o, make_float4(h[0], h[1], h[2], h[3]));
    store4(o + Lp, make_float4(m[0], m[1], m[2], m[3]));
    store4(o + 2 * Lp, make_float4(h[0], h[1], h[2], h[3]));
    store4(o + 3 * Lp, make_float4(l[0], l[1], l[2], l[3]));
    if (blocks == SPLIT_BLOCKS) {
      store4(o + 4 * Lp, make_float4(m[0], m[1], m[2], m[3]));
      store4(o + 5 * Lp, make_float4(h[0], h[1], h[2], h[3]));
    }
  }
}

}  // namespace sininn

using namespace sininn;

static int check_conv_desc(const sininn_conv_desc* d, const char* who) {
  SININN_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
  SININN_CHECK_ARG(d->in && d->wpack && d->out, "%s: null tensor pointer", who);
  SININN_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "%s: bad shape", who);
  SININN_CHECK_ARG(d->taps == 1 || d->taps == 9, "%s: taps must be 1 or 9 (got %d)", who, d->taps);
  SININN_CHECK_ARG(d->in_stride >= d->Cin && d->out_stride >= d->Cout, "%s: stride smaller than channel count", who);
  SININN_CHECK_ARG(d->rows_pad >= d->Cout && d->k_pad >= d->Cin, "%s: packed weight smaller than Cout x Cin", who);
  SININN_CHECK_ARG(!d->mask || d->mask_stride >= d->Cout, "%s: bad mask stride", who);
  return SININN_OK;
}

extern "C" {

int sininn_conv_simt(const sininn_conv_desc* d, sininn_stream_t stream) {
  int rc = check_conv_desc(d, "conv_simt");
  if (rc) return rc;
  SININN_CHECK_ARG((d->k_pad % 4) == 0, "conv_simt: k_pad must be a multiple of 4");
  SININN_CHECK_ARG(d->mask_bits == nullptr && d->bits_out == nullptr, "conv_simt: sign-bit masks are a tensor-core-path feature");
  ConvArgs a;
  a.B = d->B; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout; a.taps = d->taps;
  a.in = d->in; a.in_stride = d->in_stride; a.w = d->wpack; a.rows_pad = d->rows_pad; a.k_pad = d->k_pad;
  a.bias = d->bias; a.out = d->out; a.out_stride = d->out_stride; a.act = d->act; a.slope = d->slope;
  a.mask = d->mask; a.mask_stride = d->mask_stride; a.mask_act = d->mask_act;
  a.accumulate = d->accumulate; a.alpha = d->alpha;
  a.tiles_h = (d->H + TPX - 1) / TPX; a.tiles_w = (d->W + TPX - 1) / TPX;
  const bool f32in = d->in_dtype == SININN_F32;
  a.vec_in = ((d->in_stride % 4) == 0 && (f32in ? aligned16(d->in) : aligned8(d->in))) ? 1 : 0;
  SININN_CHECK_ARG(f32in ? aligned16(d->wpack) : aligned8(d->wpack), "conv_simt: packed weights misaligned");
  long long tiles = (long long)d->B * a.tiles_h * a.tiles_w;
  SININN_CHECK_ARG(tiles < (1ll << 31), "conv_simt: too many tiles");
  dim3 grid((unsigned)tiles, (d->Cout + BN - 1) / BN), block(256);
  cudaStream_t st = as_stream(stream);
  const bool f32out = d->out_dtype == SININN_F32;
  if (f32in && f32out) launch_k(conv_simt_kernel<float, float>, dim3(grid), dim3(block), 0, st, a);
  else if (f32in && !f32out) launch_k(conv_simt_kernel<float, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, a);
  else if (!f32in && f32out) launch_k(conv_simt_kernel<__nv_bfloat16, float>, dim3(grid), dim3(block), 0, st, a);
  else launch_k(conv_simt_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, a);
  SININN_CHECK_LAUNCH("conv_simt");
  return SININN_OK;
}

int sininn_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int taps, int mode, void* out, int out_dtype,
                            int rows_pad, int k_pad, sininn_stream_t stream) {
  SININN_CHECK_ARG(w_oihw && out && Cout > 0 && Cin > 0, "pack_conv_weight: bad arguments");
  SININN_CHECK_ARG(taps == 1 || taps == 9, "pack_conv_weight: taps must be 1 or 9");
  SININN_CHECK_ARG(mode >= 0 && mode <= 4, "pack_conv_weight: mode must be 0 (fprop), 1 (dgrad), 2 / 3 (their split forms), 4 (interleaved fprop)");
  SININN_CHECK_ARG(mode != 4 || (Cout % 2) == 0, "pack_conv_weight: the interleaved layout needs an even Cout");
  const bool split = mode == 2 || mode == 3;
  const int rows = (mode & 1) == 0 ? Cout : Cin, k = (mode & 1) == 0 ? Cin : Cout;
  SININN_CHECK_ARG(rows_pad >= rows && (!split ? k_pad >= k : ((k_pad % SPLIT_BLOCKS) == 0 && k_pad / SPLIT_BLOCKS >= k)),
                   "pack_conv_weight: padding smaller than the matrix");
  SININN_CHECK_ARG(!split || out_dtype == SININN_BF16, "pack_conv_weight: split packs are bf16");
  const long long total = (long long)taps * rows_pad * k_pad;
  long long g = (total + 255) / 256;
  if (g > (long long)sm_count() * 16) g = (long long)sm_count() * 16;
  cudaStream_t st = as_stream(stream);
  if (out_dtype == SININN_F32) launch_k(pack_weight_kernel<float>, dim3((int)g), dim3(256), 0, st, w_oihw, Cout, Cin, taps, mode, (float*)out, rows_pad, k_pad);
  else if (out_dtype == SININN_BF16) launch_k(pack_weight_kernel<__nv_bfloat16>, dim3((int)g), dim3(256), 0, st, w_oihw, Cout, Cin, taps, mode, (__nv_bfloat16*)out, rows_pad, k_pad);
  else SININN_CHECK_ARG(false, "pack_conv_weight: bad out_dtype");
  SININN_CHECK_LAUNCH("pack_conv_weight");
  return SININN_OK;
}

int sininn_pack_conv_weights_batched(const void* jobs, int njobs, int out_dtype, sininn_stream_t stream) {
  SININN_CHECK_ARG(jobs && njobs > 0, "pack_conv_weights_batched: bad arguments");
  dim3 grid(64, njobs);
  if (out_dtype == SININN_F32) launch_k(pack_weight_batched_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream), (const long long*)jobs);
  else if (out_dtype == SININN_BF16) launch_k(pack_weight_batched_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, as_stream(stream), (const long long*)jobs);
  else SININN_CHECK_ARG(false, "pack_conv_weights_batched: bad out_dtype");
  SININN_CHECK_LAUNCH("pack_conv_weights_batched");
  return SININN_OK;
}

int sininn_split_bf16(const float* in, int in_stride, long long npix, int L, float scale, void* out, int Lp, int blocks,
                      sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && npix > 0 && L > 0, "split_bf16: bad arguments");
  SININN_CHECK_ARG(blocks == 4 || blocks == SPLIT_BLOCKS, "split_bf16: 4 or 6 channel blocks");
  SININN_CHECK_ARG(Lp >= L && (Lp % 8) == 0 && aligned8(out), "split_bf16: block width must be a multiple of 8 channels >= L");
  const long long total = npix * (Lp / 4);
  long long g = (total + 255) / 256;
  if (g > (long long)sm_count() * 32) g = (long long)sm_count() * 32;
  launch_k(split_slice_kernel, dim3((int)g), dim3(256), 0, as_stream(stream), in, in_stride, npix, L, Lp, scale, (__nv_bfloat16*)out, blocks);
  SININN_CHECK_LAUNCH("split_bf16");
  return SININN_OK;
}

int sininn_wgrad_simt(const sininn_wgrad_desc* d, sininn_stream_t stream) {
  SININN_CHECK_ARG(d && d->x && d->dy && d->dw, "wgrad_simt: null pointer");
  SININN_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "wgrad_simt: bad shape");
  SININN_CHECK_ARG(d->taps == 1 || d->taps == 9, "wgrad_simt: taps must be 1 or 9");
  const int splits = wgrad_splits(d);
  const size_t need = (size_t)splits * d->taps * d->Cout * d->Cin * sizeof(float);
  if (!d->workspace || d->workspace_bytes < need) {
    set_error("wgrad_simt: workspace too small (%zu < %zu)", d->workspace_bytes, need);
    return SININN_EWORKSPACE;
  }
  WgradArgs a;
  a.B = d->B; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout; a.taps = d->taps;
  a.x = d->x; a.x_stride = d->x_stride; a.dy = d->dy; a.dy_stride = d->dy_stride;
  a.partial = reinterpret_cast<float*>(d->workspace);
  a.npix = (long long)d->B * d->H * d->W;
  a.pix_per_split = (a.npix + splits - 1) / splits;
  a.ci_tiles = (d->Cin + BN - 1) / BN;
  const bool xf = d->x_dtype == SININN_F32, df = d->dy_dtype == SININN_F32;
  a.vec_x = ((d->x_stride % 4) == 0 && (xf ? aligned16(d->x) : aligned8(d->x))) ? 1 : 0;
  a.vec_dy = ((d->dy_stride % 4) == 0 && (df ? aligned16(d->dy) : aligned8(d->dy))) ? 1 : 0;
  dim3 grid((d->Cout + BM - 1) / BM, a.ci_tiles * d->taps, splits), block(256);
  cudaStream_t st = as_stream(stream);
  if (xf && df) launch_k(wgrad_simt_kernel<float, float>, dim3(grid), dim3(block), 0, st, a);
  else if (xf && !df) launch_k(wgrad_simt_kernel<float, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, a);
  else if (!xf && df) launch_k(wgrad_simt_kernel<__nv_bfloat16, float>, dim3(grid), dim3(block), 0, st, a);
  else launch_k(wgrad_simt_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, a);
  const long long per = (long long)d->taps * d->Cout * d->Cin;
  long long g = (per + 255) / 256;
  if (g > (long long)sm_count() * 8) g = (long long)sm_count() * 8;
  launch_k(wgrad_reduce_kernel, dim3((int)g), dim3(256), 0, st, a.partial, splits, d->taps, d->Cout, d->Cin, d->dw, d->accumulate);
  SININN_CHECK_LAUNCH("wgrad_simt");
  return SININN_OK;
}

}  // extern "C"

// shared with conv_tc.cu
namespace sininn {
int wgrad_simt_splits(const sininn_wgrad_desc* d) { return wgrad_splits(d); }
}
