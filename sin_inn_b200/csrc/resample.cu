// Squeeze (i-RevNet order) and Haar down/up-sampling, layout changes, channel permutation.
// All HBM-bound: one read + one write of the tensor, 128-bit accesses where alignment allows.
//
// Reference semantics:
//   squeeze  FrEIA IRevNetDownsampling (call sites /root/reference/archs.py:28-31,35-38)
//            out[b,(dy*2+dx)*C+c,i,j] = in[b,c,2i+dy,2j+dx]
//   Haar     HaarDownsampling.forward /root/reference/archs.py:183-199, patterns :167-176,
//            band-major channel order out[k*C+c] (:188-190)
#include "common.cuh"

namespace sininn {

// a=(0,0) b=(0,1) c=(1,0) d=(1,1) of one 2x2 block  ->  four output planes
template <bool HAAR>
__device__ __forceinline__ void fwd4(float a, float b, float c, float d, float scale, float o[4]) {
  if (HAAR) {
    o[0] = (a + b + c + d) * scale;
    o[1] = (a - b + c - d) * scale;
    o[2] = (a + b - c - d) * scale;
    o[3] = (a - b - c + d) * scale;
  } else {
    o[0] = a; o[1] = b; o[2] = c; o[3] = d;
  }
}
// inverse of the reference forward is the transpose of the same +-1 patterns (H*H^T = 4I)
template <bool HAAR>
__device__ __forceinline__ void inv4(const float o[4], float scale, float& a, float& b, float& c, float& d) {
  if (HAAR) {
    a = (o[0] + o[1] + o[2] + o[3]) * scale;
    b = (o[0] - o[1] + o[2] - o[3]) * scale;
    c = (o[0] + o[1] - o[2] - o[3]) * scale;
    d = (o[0] - o[1] - o[2] + o[3]) * scale;
  } else {
    a = o[0]; b = o[1]; c = o[2]; d = o[3];
  }
}

template <int N> struct VecT;
template <> struct VecT<1> { typedef float type; };
template <> struct VecT<2> { typedef float2 type; };
template <> struct VecT<4> { typedef float4 type; };

// ---------------------------------------------------------------- NCHW <-> NCHW
// One thread = VEC output columns of one (plane, output row): reads 2 rows x 2*VEC floats,
// writes VEC floats to each of 4 planes.  Consecutive threads walk the row => coalesced both ways.
template <int VEC, bool HAAR>
__global__ void __launch_bounds__(256) resample_nchw_fwd_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                int C, int H, int W, float scale, long long total) {
  pdl_wait();
  pdl_trigger();
  const int Ho = H >> 1, Wo = W >> 1, Wv = Wo / VEC;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int jv;
    long long r;
    split_index(idx, Wv, r, jv);
    int i, c;
    long long plane, b;                // plane = b*C + c
    split_index(r, Ho, plane, i);
    split_index(plane, C, b, c);
    const float* r0 = in + (plane * H + 2 * i) * (long long)W + 2 * jv * VEC;
    const float* r1 = r0 + W;
    __align__(16) float x0[2 * VEC], x1[2 * VEC];
    typedef typename VecT<VEC == 1 ? 2 : 4>::type LT;   // VEC=1: float2, VEC>=2: float4 chunks
    constexpr int LN = (VEC == 1 ? 2 : 4);
#pragma unroll
    for (int q = 0; q < 2 * VEC / LN; ++q) {
      *reinterpret_cast<LT*>(&x0[q * LN]) = __ldcs(reinterpret_cast<const LT*>(r0) + q);
      *reinterpret_cast<LT*>(&x1[q * LN]) = __ldcs(reinterpret_cast<const LT*>(r1) + q);
    }
    __align__(16) float o[4][VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float t[4];
      fwd4<HAAR>(x0[2 * v], x0[2 * v + 1], x1[2 * v], x1[2 * v + 1], scale, t);
      o[0][v] = t[0]; o[1][v] = t[1]; o[2][v] = t[2]; o[3][v] = t[3];
    }
    typedef typename VecT<VEC>::type ST;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float* dst = out + (((b * 4 + k) * C + c) * Ho + i) * (long long)Wo + jv * VEC;
      __stcs(reinterpret_cast<ST*>(dst), *reinterpret_cast<ST*>(&o[k][0]));
    }
  }
}

template <int VEC, bool HAAR>
__global__ void __launch_bounds__(256) resample_nchw_inv_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                int C, int H, int W, float scale, long long total) {
  pdl_wait();
  pdl_trigger();
  const int Ho = H >> 1, Wo = W >> 1, Wv = Wo / VEC;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int jv;
    long long r;
    split_index(idx, Wv, r, jv);
    int i, c;
    long long plane, b;
    split_index(r, Ho, plane, i);
    split_index(plane, C, b, c);
    typedef typename VecT<VEC>::type ST;
    __align__(16) float o[4][VEC];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* src = in + (((b * 4 + k) * C + c) * Ho + i) * (long long)Wo + jv * VEC;
      *reinterpret_cast<ST*>(&o[k][0]) = __ldcs(reinterpret_cast<const ST*>(src));
    }
    __align__(16) float x0[2 * VEC], x1[2 * VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float t[4] = {o[0][v], o[1][v], o[2][v], o[3][v]};
      inv4<HAAR>(t, scale, x0[2 * v], x0[2 * v + 1], x1[2 * v], x1[2 * v + 1]);
    }
    float* r0 = out + (plane * H + 2 * i) * (long long)W + 2 * jv * VEC;
    float* r1 = r0 + W;
    typedef typename VecT<VEC == 1 ? 2 : 4>::type LT;
    constexpr int LN = (VEC == 1 ? 2 : 4);
#pragma unroll
    for (int q = 0; q < 2 * VEC / LN; ++q) {
      __stcs(reinterpret_cast<LT*>(r0) + q, *reinterpret_cast<LT*>(&x0[q * LN]));
      __stcs(reinterpret_cast<LT*>(r1) + q, *reinterpret_cast<LT*>(&x1[q * LN]));
    }
  }
}

// ---------------------------------------------------------------- NHWC <-> NHWC
// full-res [B,H,W,C] <-> [B,H/2,W/2,4C]; one thread = VEC channels of one low-res pixel.
template <int VEC, bool HAAR, bool REV>
__global__ void __launch_bounds__(256) resample_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            int C, int H, int W, float scale, long long total) {
  pdl_wait();
  pdl_trigger();
  const int Ho = H >> 1, Wo = W >> 1, Cv = C / VEC;
  typedef typename VecT<VEC>::type VT;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int cv;
    long long r;
    split_index(idx, Cv, r, cv);
    int j, i;
    long long r2, b;
    split_index(r, Wo, r2, j);
    split_index(r2, Ho, b, i);
    int c = cv * VEC;
    long long hi = ((b * H + 2 * i) * (long long)W + 2 * j) * C + c;   // full-res (2i,2j)
    long long lo = ((b * Ho + i) * (long long)Wo + j) * 4 * C + c;     // low-res pixel, band 0
    __align__(16) float v[4][VEC];
    if (!REV) {
      *reinterpret_cast<VT*>(v[0]) = *reinterpret_cast<const VT*>(in + hi);
      *reinterpret_cast<VT*>(v[1]) = *reinterpret_cast<const VT*>(in + hi + C);
      *reinterpret_cast<VT*>(v[2]) = *reinterpret_cast<const VT*>(in + hi + (long long)W * C);
      *reinterpret_cast<VT*>(v[3]) = *reinterpret_cast<const VT*>(in + hi + (long long)W * C + C);
      __align__(16) float o[4][VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float t[4];
        fwd4<HAAR>(v[0][e], v[1][e], v[2][e], v[3][e], scale, t);
        o[0][e] = t[0]; o[1][e] = t[1]; o[2][e] = t[2]; o[3][e] = t[3];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<VT*>(out + lo + (long long)k * C) = *reinterpret_cast<VT*>(o[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) *reinterpret_cast<VT*>(v[k]) = *reinterpret_cast<const VT*>(in + lo + (long long)k * C);
      __align__(16) float o[4][VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float t[4] = {v[0][e], v[1][e], v[2][e], v[3][e]};
        inv4<HAAR>(t, scale, o[0][e], o[1][e], o[2][e], o[3][e]);
      }
      *reinterpret_cast<VT*>(out + hi) = *reinterpret_cast<VT*>(o[0]);
      *reinterpret_cast<VT*>(out + hi + C) = *reinterpret_cast<VT*>(o[1]);
      *reinterpret_cast<VT*>(out + hi + (long long)W * C) = *reinterpret_cast<VT*>(o[2]);
      *reinterpret_cast<VT*>(out + hi + (long long)W * C + C) = *reinterpret_cast<VT*>(o[3]);
    }
  }
}

// ---------------------------------------------------------------- NCHW <-> NHWC (32x32 smem transpose)
// TO_NHWC: in [B][C][HW] -> out [B][HW][C], out channel i <- in channel map[i]
// else   : in [B][HW][C] -> out [B][C][HW], out channel i <- in channel map[i]
template <bool TO_NHWC>
__global__ void __launch_bounds__(256) layout_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW,
                                                     const int32_t* __restrict__ map, __nv_bfloat16* __restrict__ bf,
                                                     int bc0, int bc1) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (TO_NHWC) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int c = c0 + ty + 8 * r, p = p0 + tx;
      if (c < C && p < HW) {
        int cs = map ? map[c] : c;
        tile[ty + 8 * r][tx] = __ldcs(in + (b * C + cs) * (long long)HW + p);
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int p = p0 + ty + 8 * r, c = c0 + tx;
      if (c < C && p < HW) {
        float v = tile[tx][ty + 8 * r];
        out[(b * HW + p) * (long long)C + c] = v;
        if (bf != nullptr && c >= bc0 && c < bc1) bf[(b * HW + p) * (long long)(bc1 - bc0) + (c - bc0)] = __float2bfloat16_rn(v);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int p = p0 + ty + 8 * r, c = c0 + tx;
      if (c < C && p < HW) {
        int cs = map ? map[c] : c;
        tile[ty + 8 * r][tx] = in[(b * HW + p) * (long long)C + cs];
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int c = c0 + ty + 8 * r, p = p0 + tx;
      if (c < C && p < HW) __stcs(out + (b * C + c) * (long long)HW + p, tile[tx][ty + 8 * r]);
    }
  }
}

// ---------------------------------------------------------------- two squeezes + NCHW -> NHWC in one pass (and back)
// The SRF graph starts with two IRevNetDownsampling nodes (/root/reference/archs.py:28-38) before its first coupling block; as
// three kernels (squeeze, squeeze, layout change) the input is read and written three times.  Composed index map:
//   out[b][i][j][k2 * 4 C0 + k1 * C0 + c] = in[b][c][4 i + 2 dy(k2) + dy(k1)][4 j + 2 dx(k2) + dx(k1)],   dy(k) = k >> 1, dx(k) = k & 1.
// One block = one output row segment of SQ_P pixels of one sample: 4 input rows x C0 channels are read as float4 (the four
// values are the (dx2, dx1) positions of one output pixel), staged as [pixel][16 C0] in shared memory and written as one
// contiguous run (TO_NHWC), or the other way round.
constexpr int SQ_P = 64;
template <bool TO_NHWC>
__global__ void __launch_bounds__(256) squeeze2_layout_kernel(const float* __restrict__ in, float* __restrict__ out, int C0, int H, int W,
                                                              __nv_bfloat16* __restrict__ bf, int bc0, int bc1) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sq_tile[];                   // [SQ_P][C + 1], C = 16 * C0
  const int C = 16 * C0, pitch = C + 1;
  const int Ho = H >> 2, Wo = W >> 2;
  const int j0 = blockIdx.x * SQ_P, i = blockIdx.y;
  const long long b = blockIdx.z;
  const int np = min(SQ_P, Wo - j0);
  const int nvec = C0 * 4 * np;                        // float4 pieces of the full-resolution side: (c, row r, pixel jj)
  const float* full = TO_NHWC ? in : out;              // (address arithmetic only)
  (void)full;
  if (TO_NHWC) {
    for (int e = threadIdx.x; e < nvec; e += 256) {
      const int jj = e % np, cr = e / np, r = cr & 3, c = cr >> 2;
      const float4 v = __ldcs(reinterpret_cast<const float4*>(in + ((b * C0 + c) * H + 4 * i + r) * (long long)W + 4 * (j0 + jj)));
      const int dy2 = r >> 1, dy1 = r & 1;
      float* t = sq_tile + jj * pitch + c;
      // x offset 0..3 = 2 dx2 + dx1
      t[((dy2 * 2 + 0) * 4 + (dy1 * 2 + 0)) * C0] = v.x;
      t[((dy2 * 2 + 0) * 4 + (dy1 * 2 + 1)) * C0] = v.y;
      t[((dy2 * 2 + 1) * 4 + (dy1 * 2 + 0)) * C0] = v.z;
      t[((dy2 * 2 + 1) * 4 + (dy1 * 2 + 1)) * C0] = v.w;
    }
    __syncthreads();
    const long long pix0 = (b * Ho + i) * (long long)Wo + j0;
    float* o = out + pix0 * C;
    const int L = bc1 - bc0;
    for (int e = threadIdx.x; e < np * C; e += 256) {
      const int jj = e / C, ch = e - jj * C;
      const float v = sq_tile[jj * pitch + ch];
      o[e] = v;
      if (bf != nullptr && ch >= bc0 && ch < bc1) bf[(pix0 + jj) * L + (ch - bc0)] = __float2bfloat16_rn(v);
    }
  } else {
    const float* src = in + ((b * Ho + i) * (long long)Wo + j0) * C;
    for (int e = threadIdx.x; e < np * C; e += 256) {
      const int jj = e / C, ch = e - jj * C;
      sq_tile[jj * pitch + ch] = src[e];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nvec; e += 256) {
      const int jj = e % np, cr = e / np, r = cr & 3, c = cr >> 2;
      const int dy2 = r >> 1, dy1 = r & 1;
      const float* t = sq_tile + jj * pitch + c;
      float4 v;
      v.x = t[((dy2 * 2 + 0) * 4 + (dy1 * 2 + 0)) * C0];
      v.y = t[((dy2 * 2 + 0) * 4 + (dy1 * 2 + 1)) * C0];
      v.z = t[((dy2 * 2 + 1) * 4 + (dy1 * 2 + 0)) * C0];
      v.w = t[((dy2 * 2 + 1) * 4 + (dy1 * 2 + 1)) * C0];
      __stcs(reinterpret_cast<float4*>(out + ((b * C0 + c) * H + 4 * i + r) * (long long)W + 4 * (j0 + jj)), v);
    }
  }
}

// out[p][i] = in[p][map[i]]  (+ optional compact bf16 copy of out[:, bc0:bc1])
template <int VEC>
__global__ void __launch_bounds__(256) permute_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, long long npix,
                                                           int C, const int32_t* __restrict__ map,
                                                           __nv_bfloat16* __restrict__ bf, int bc0, int bc1) {
  pdl_wait();
  pdl_trigger();
  const int Cv = C / VEC;
  const long long total = npix * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int cv;
    long long p;
    split_index(idx, Cv, p, cv);
    int c = cv * VEC;
    float v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = in[p * C + __ldg(map + c + e)];
    if (VEC == 4) {
      store4(out + p * C + c, make_float4(v[0], v[1], v[2], v[3]));
      if (bf != nullptr && c >= bc0 && c < bc1)   // bc0, bc1 multiples of 4 on this path
        store4(bf + p * (long long)(bc1 - bc0) + (c - bc0), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        out[p * C + c + e] = v[e];
        if (bf != nullptr && c + e >= bc0 && c + e < bc1)
          bf[p * (long long)(bc1 - bc0) + (c + e - bc0)] = __float2bfloat16_rn(v[e]);
      }
    }
  }
}

// the backward pass undoes a permutation on the trunk AND its gradient with the same map: one launch for both
__global__ void __launch_bounds__(256) permute_nhwc_pair_kernel(const float* __restrict__ in_a, float* __restrict__ out_a,
                                                                const float* __restrict__ in_b, float* __restrict__ out_b,
                                                                long long npix, int C, const int32_t* __restrict__ map,
                                                                __nv_bfloat16* __restrict__ bf, int bc0, int bc1) {
  pdl_wait();
  pdl_trigger();
  const int Cv = C / 4;
  const long long total = npix * Cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Cv, p, c);
    c *= 4;
    const int m0 = __ldg(map + c), m1 = __ldg(map + c + 1), m2 = __ldg(map + c + 2), m3 = __ldg(map + c + 3);
    const float* ra = in_a + p * C;
    const float* rb = in_b + p * C;
    const float4 va = make_float4(ra[m0], ra[m1], ra[m2], ra[m3]);
    store4(out_a + p * C + c, va);
    store4(out_b + p * C + c, make_float4(rb[m0], rb[m1], rb[m2], rb[m3]));
    if (bf != nullptr && c >= bc0 && c < bc1)      // compact bf16 copy of out_a[:, bc0:bc1] (bc0, bc1 multiples of 4)
      store4(bf + p * (long long)(bc1 - bc0) + (c - bc0), va);
  }
}

// fp32 NCHW in [0, 1] -> uint8 HWC: what the reference's inference loop does per image on the host
// (lit_wrapper.py:117-121: transforms.ToPILImage() = pic.mul(255).byte(), then channels-last for PIL)
__global__ void __launch_bounds__(256) quantize_u8_hwc_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int C,
                                                              long long HW, long long total_pix) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_pix; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float* src = in + b * C * HW + p;
    uint8_t* dst = out + i * C;
    for (int c = 0; c < C; ++c) {
      float v = src[c * HW];
      v = fminf(fmaxf(v, 0.f), 1.f) * 255.f;          // in-range values: identical to mul(255).byte() (truncation)
      dst[c] = (uint8_t)(int)v;
    }
  }
}

// uint8 video frames resident on the GPU -> one fp32 NCHW training batch: for every sample the frames
// centre-win .. centre+win of a [T][H][W][C] uint8 video, cropped to a ph x pw patch, concatenated along channels and
// divided by 255 (data.py:31-45: imread of 2*lr_window+1 files, concatenate(axis=-1), transpose(-1, 0, 1), / 255.)
__global__ void __launch_bounds__(256) gather_windows_u8_kernel(const uint8_t* __restrict__ video, int T, int H, int W, int C,
                                                                const int32_t* __restrict__ centers, int win, int y0, int x0,
                                                                int ph, int pw, float* __restrict__ out, long long total,
                                                                const int32_t* __restrict__ crops) {
  pdl_wait();
  pdl_trigger();
  const int F = 2 * win + 1;
  const long long plane = (long long)ph * pw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % pw);
    long long r = i / pw;
    const int y = (int)(r % ph); r /= ph;
    const int f = (int)(r % F);
    const int b = (int)(r / F);
    int t = __ldg(centers + b) - win + f;
    t = t < 0 ? 0 : (t >= T ? T - 1 : t);                  // (the reference never indexes outside the clip)
    const int yy = crops ? __ldg(crops + 2 * b) : y0, xx = crops ? __ldg(crops + 2 * b + 1) : x0;   // per-sample patch origin
    const uint8_t* src = video + (((long long)t * H + (yy + y)) * W + (xx + x)) * C;
    float* dst = out + ((long long)b * F * C + (long long)f * C) * plane + (long long)y * pw + x;
    for (int c = 0; c < C; ++c) dst[c * plane] = __fdiv_rn((float)src[c], 255.0f);      // bit-exact with the reference's "/ 255."
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = (long long)sm_count() * 32;   // grid-stride beyond ~32 CTAs per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace sininn

using namespace sininn;

extern "C" {

int sininn_resample_nchw(const float* in, float* out, int B, int C, int H, int W, int mode, int rev, float scale,
                         sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out, "resample_nchw: null pointer");
  SININN_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "resample_nchw: bad shape");
  SININN_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "resample_nchw: H and W must be even (got %dx%d)", H, W);
  SININN_CHECK_ARG(mode == 0 || mode == 1, "resample_nchw: mode must be 0 (squeeze) or 1 (haar)");
  cudaStream_t st = as_stream(stream);
  const int Wo = W / 2;
  int vec = 1;
  if ((Wo % 4) == 0 && aligned16(in) && aligned16(out)) vec = 4;
  else if ((Wo % 2) == 0 && aligned16(in) && aligned16(out)) vec = 2;
  else SININN_CHECK_ARG(aligned8(in) && aligned8(out), "resample_nchw: pointers must be 8-byte aligned");
  const long long total = (long long)B * C * (H / 2) * (Wo / vec);
  const int block = 256, grid = grid_for(total, block);
#define LAUNCH(V, HA)                                                                                   \
  do {                                                                                                  \
    if (!rev) launch_k(resample_nchw_fwd_kernel<V, HA>, dim3(grid), dim3(block), 0, st, in, out, C, H, W, scale, total);  \
    else launch_k(resample_nchw_inv_kernel<V, HA>, dim3(grid), dim3(block), 0, st, in, out, C, H, W, scale, total);       \
  } while (0)
  if (mode == 1) {
    if (vec == 4) LAUNCH(4, true); else if (vec == 2) LAUNCH(2, true); else LAUNCH(1, true);
  } else {
    if (vec == 4) LAUNCH(4, false); else if (vec == 2) LAUNCH(2, false); else LAUNCH(1, false);
  }
#undef LAUNCH
  SININN_CHECK_LAUNCH("resample_nchw");
  return SININN_OK;
}

int sininn_resample_nhwc(const float* in, float* out, int B, int C, int H, int W, int mode, int rev, float scale,
                         sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out, "resample_nhwc: null pointer");
  SININN_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "resample_nhwc: bad shape");
  SININN_CHECK_ARG((H % 2) == 0 && (W % 2) == 0, "resample_nhwc: H and W must be even (got %dx%d)", H, W);
  SININN_CHECK_ARG(mode == 0 || mode == 1, "resample_nhwc: mode must be 0 (squeeze) or 1 (haar)");
  cudaStream_t st = as_stream(stream);
  const int vec = ((C % 4) == 0 && aligned16(in) && aligned16(out)) ? 4 : 1;
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / vec);
  const int block = 256, grid = grid_for(total, block);
#define LAUNCH(V, HA, RV) launch_k(resample_nhwc_kernel<V, HA, RV>, dim3(grid), dim3(block), 0, st, in, out, C, H, W, scale, total)
  if (vec == 4) {
    if (mode == 1) { if (rev) LAUNCH(4, true, true); else LAUNCH(4, true, false); }
    else           { if (rev) LAUNCH(4, false, true); else LAUNCH(4, false, false); }
  } else {
    if (mode == 1) { if (rev) LAUNCH(1, true, true); else LAUNCH(1, true, false); }
    else           { if (rev) LAUNCH(1, false, true); else LAUNCH(1, false, false); }
  }
#undef LAUNCH
  SININN_CHECK_LAUNCH("resample_nhwc");
  return SININN_OK;
}

int sininn_nchw_to_nhwc(const float* in, float* out, int B, int C, int HW, const int32_t* chan_map, void* bf16_out,
                        int c0, int c1, sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && B > 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad arguments");
  SININN_CHECK_ARG(B <= 65535, "nchw_to_nhwc: batch too large for grid.z");
  if (bf16_out) SININN_CHECK_ARG(0 <= c0 && c0 < c1 && c1 <= C, "nchw_to_nhwc: bad bf16 channel range");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_k(layout_kernel<true>, dim3(grid), dim3(block), 0, as_stream(stream), in, out, C, HW, chan_map,
                                                             reinterpret_cast<__nv_bfloat16*>(bf16_out), c0, c1);
  SININN_CHECK_LAUNCH("nchw_to_nhwc");
  return SININN_OK;
}

int sininn_nhwc_to_nchw(const float* in, float* out, int B, int C, int HW, const int32_t* chan_map,
                        sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && B > 0 && C > 0 && HW > 0, "nhwc_to_nchw: bad arguments");
  SININN_CHECK_ARG(B <= 65535, "nhwc_to_nchw: batch too large for grid.z");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_k(layout_kernel<false>, dim3(grid), dim3(block), 0, as_stream(stream), in, out, C, HW, chan_map, nullptr, 0, 0);
  SININN_CHECK_LAUNCH("nhwc_to_nchw");
  return SININN_OK;
}

int sininn_squeeze2_to_nhwc(const float* in, float* out, int B, int C0, int H, int W, void* bf16_out, int c0, int c1,
                            sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && B > 0 && C0 > 0 && H > 0 && W > 0, "squeeze2_to_nhwc: bad arguments");
  SININN_CHECK_ARG((H % 4) == 0 && (W % 4) == 0 && aligned16(in), "squeeze2_to_nhwc: H and W must be multiples of 4 and the input 16-byte aligned (got %dx%d)", H, W);
  SININN_CHECK_ARG(B <= 65535 && H / 4 <= 65535 && 16 * C0 <= 128, "squeeze2_to_nhwc: batch / height / channels out of range");
  if (bf16_out) SININN_CHECK_ARG(0 <= c0 && c0 < c1 && c1 <= 16 * C0, "squeeze2_to_nhwc: bad bf16 channel range");
  const int Wo = W / 4;
  dim3 grid((Wo + SQ_P - 1) / SQ_P, H / 4, B);
  launch_k(squeeze2_layout_kernel<true>, grid, dim3(256), (size_t)SQ_P * (16 * C0 + 1) * sizeof(float), as_stream(stream), in, out, C0, H, W,
           reinterpret_cast<__nv_bfloat16*>(bf16_out), c0, c1);
  SININN_CHECK_LAUNCH("squeeze2_to_nhwc");
  return SININN_OK;
}

int sininn_nhwc_to_unsqueeze2(const float* in, float* out, int B, int C0, int H, int W, sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && B > 0 && C0 > 0 && H > 0 && W > 0, "nhwc_to_unsqueeze2: bad arguments");
  SININN_CHECK_ARG((H % 4) == 0 && (W % 4) == 0 && aligned16(out), "nhwc_to_unsqueeze2: H and W must be multiples of 4 and the output 16-byte aligned (got %dx%d)", H, W);
  SININN_CHECK_ARG(B <= 65535 && H / 4 <= 65535 && 16 * C0 <= 128, "nhwc_to_unsqueeze2: batch / height / channels out of range");
  const int Wo = W / 4;
  dim3 grid((Wo + SQ_P - 1) / SQ_P, H / 4, B);
  launch_k(squeeze2_layout_kernel<false>, grid, dim3(256), (size_t)SQ_P * (16 * C0 + 1) * sizeof(float), as_stream(stream), in, out, C0, H, W,
           static_cast<__nv_bfloat16*>(nullptr), 0, 0);
  SININN_CHECK_LAUNCH("nhwc_to_unsqueeze2");
  return SININN_OK;
}

int sininn_permute_nhwc(const float* in, float* out, long long npix, int C, const int32_t* chan_map, void* bf16_out,
                        int c0, int c1, sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && chan_map && npix > 0 && C > 0, "permute_nhwc: bad arguments");
  SININN_CHECK_ARG(in != out, "permute_nhwc: cannot run in place");
  if (bf16_out) SININN_CHECK_ARG(0 <= c0 && c0 < c1 && c1 <= C, "permute_nhwc: bad bf16 channel range");
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(bf16_out);
  bool v4 = (C % 4) == 0 && aligned16(out) && (!bf || ((c0 % 4) == 0 && (c1 % 4) == 0 && aligned8(bf)));
  const long long total = npix * (v4 ? C / 4 : C);
  const int block = 256, grid = grid_for(total, block);
  if (v4) launch_k(permute_nhwc_kernel<4>, dim3(grid), dim3(block), 0, as_stream(stream), in, out, npix, C, chan_map, bf, c0, c1);
  else launch_k(permute_nhwc_kernel<1>, dim3(grid), dim3(block), 0, as_stream(stream), in, out, npix, C, chan_map, bf, c0, c1);
  SININN_CHECK_LAUNCH("permute_nhwc");
  return SININN_OK;
}

int sininn_gather_windows_u8(const uint8_t* video, int T, int H, int W, int C, const int32_t* centers, int B, int win,
                             int y0, int x0, int ph, int pw, float* out, sininn_stream_t stream) {
  SININN_CHECK_ARG(video && centers && out && T > 0 && H > 0 && W > 0 && C > 0 && B > 0 && win >= 0, "gather_windows_u8: bad arguments");
  SININN_CHECK_ARG(y0 >= 0 && x0 >= 0 && ph > 0 && pw > 0 && y0 + ph <= H && x0 + pw <= W, "gather_windows_u8: crop outside the frame");
  const long long total = (long long)B * (2 * win + 1) * ph * pw;
  launch_k(gather_windows_u8_kernel, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), video, T, H, W, C, centers, win, y0, x0,
           ph, pw, out, total, (const int32_t*)nullptr);
  SININN_CHECK_LAUNCH("gather_windows_u8");
  return SININN_OK;
}

int sininn_gather_windows_u8_crops(const uint8_t* video, int T, int H, int W, int C, const int32_t* centers, int B, int win,
                                   const int32_t* crops_yx, int ph, int pw, float* out, sininn_stream_t stream) {
  SININN_CHECK_ARG(video && centers && crops_yx && out && T > 0 && H > 0 && W > 0 && C > 0 && B > 0 && win >= 0,
                   "gather_windows_u8_crops: bad arguments");
  SININN_CHECK_ARG(ph > 0 && pw > 0 && ph <= H && pw <= W, "gather_windows_u8_crops: patch larger than the frame");
  const long long total = (long long)B * (2 * win + 1) * ph * pw;
  launch_k(gather_windows_u8_kernel, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), video, T, H, W, C, centers, win, 0, 0,
           ph, pw, out, total, crops_yx);
  SININN_CHECK_LAUNCH("gather_windows_u8_crops");
  return SININN_OK;
}

int sininn_quantize_u8_hwc(const float* in, uint8_t* out, int B, int C, int H, int W, sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0, "quantize_u8_hwc: bad arguments");
  const long long HW = (long long)H * W, total = (long long)B * HW;
  launch_k(quantize_u8_hwc_kernel, dim3(grid_for(total, 256)), dim3(256), 0, as_stream(stream), in, out, C, HW, total);
  SININN_CHECK_LAUNCH("quantize_u8_hwc");
  return SININN_OK;
}

int sininn_permute_nhwc_pair(const float* in_a, float* out_a, const float* in_b, float* out_b, long long npix, int C,
                             const int32_t* chan_map, void* bf16_out_a, int c0, int c1, sininn_stream_t stream) {
  SININN_CHECK_ARG(in_a && out_a && in_b && out_b && chan_map && npix > 0 && C > 0, "permute_nhwc_pair: bad arguments");
  SININN_CHECK_ARG(in_a != out_a && in_b != out_b, "permute_nhwc_pair: cannot run in place");
  SININN_CHECK_ARG((C % 4) == 0 && aligned16(out_a) && aligned16(out_b), "permute_nhwc_pair: needs C %% 4 == 0 and 16-byte aligned outputs");
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(bf16_out_a);
  if (bf) SININN_CHECK_ARG(0 <= c0 && c0 < c1 && c1 <= C && (c0 % 4) == 0 && (c1 % 4) == 0 && aligned8(bf), "permute_nhwc_pair: bad bf16 channel range");
  const long long total = npix * (C / 4);
  const int block = 256, grid = grid_for(total, block);
  launch_k(permute_nhwc_pair_kernel, dim3(grid), dim3(block), 0, as_stream(stream), in_a, out_a, in_b, out_b, npix, C, chan_map, bf, c0, c1);
  SININN_CHECK_LAUNCH("permute_nhwc_pair");
  return SININN_OK;
}

}  // extern "C"
