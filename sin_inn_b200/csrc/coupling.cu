// Affine-coupling elementwise kernels (forward, inverse, backward-from-output) and small
// bandwidth-bound helpers around the subnets.  All operate on channels-last [npix][L] slices
// addressed with explicit pixel strides, 128-bit accesses when L and the strides allow.
//
// Reference semantics:
//   GLOW  y = exp(g(s))*x + t / x = (y-t)/exp(g(s)), g(s)=clamp*0.636*atan(s/clamp)
//         (FrEIA GLOWCouplingBlock pre-v0.2; call site /root/reference/archs.py:61-64)
//   IRN   y2 = x2*exp(s) + G(y1), s = clamp*(2*sigmoid(H(y1))-1)   (/root/reference/archs.py:152-158)
// Backward-from-output equations: SURVEY.md section 8a.
#include "common.cuh"

namespace sininn {

template <int VEC>
struct Pack {
  float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Pack<VEC> ldv(const float* p) {
  Pack<VEC> r;
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = *p;
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const Pack<VEC>& r) {
  if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  else *p = r.v[0];
}
template <int VEC>
__device__ __forceinline__ void stv(__nv_bfloat16* p, const Pack<VEC>& r) {
  if constexpr (VEC == 4) store4(p, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  else *p = __float2bfloat16_rn(r.v[0]);
}

template <int VEC, bool FAST>
__global__ void __launch_bounds__(256) coupling_apply_kernel(float* __restrict__ u, int u_stride, const float* __restrict__ s,
                                                             int s_stride, const float* __restrict__ t, int t_stride,
                                                             long long npix, int L, int kind, float clamp, int inverse,
                                                             __nv_bfloat16* __restrict__ ubf) {
  pdl_wait();
  pdl_trigger();
  const int Lv = L / VEC;
  const long long total = npix * Lv;
  const float inv_clamp = 1.0f / clamp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Lv, p, c);
    c *= VEC;
    Pack<VEC> x = ldv<VEC>(u + p * u_stride + c);
    Pack<VEC> sv = ldv<VEC>(s + p * s_stride + c);
    Pack<VEC> tv = ldv<VEC>(t + p * t_stride + c);
    Pack<VEC> y;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float ex, dg;
      if (FAST) {
        scale_fast(kind, clamp, inv_clamp, sv.v[e], ex, dg);
        y.v[e] = inverse ? __fdividef(x.v[e] - tv.v[e], ex) : fmaf(ex, x.v[e], tv.v[e]);
      } else {
        float g;
        log_scale(kind, clamp, sv.v[e], g, dg);
        ex = expf(g);
        y.v[e] = inverse ? (x.v[e] - tv.v[e]) / ex : ex * x.v[e] + tv.v[e];
      }
    }
    stv<VEC>(u + p * u_stride + c, y);
    if (ubf != nullptr) stv<VEC>(ubf + p * L + c, y);
  }
}

template <int VEC, typename TO, bool FAST>
__global__ void __launch_bounds__(256) coupling_bwd_kernel(float* __restrict__ u, int u_stride, float* __restrict__ du, int du_stride,
                                                           const float* __restrict__ s, int s_stride, const float* __restrict__ t,
                                                           int t_stride, long long npix, int L, int kind, float clamp, int inverse,
                                                           TO* __restrict__ ds, int ds_stride, TO* __restrict__ dt, int dt_stride,
                                                           __nv_bfloat16* __restrict__ xbf) {
  pdl_wait();
  pdl_trigger();
  const int Lv = L / VEC;
  const long long total = npix * Lv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Lv, p, c);
    c *= VEC;
    Pack<VEC> y = ldv<VEC>(u + p * u_stride + c);
    Pack<VEC> dy = ldv<VEC>(du + p * du_stride + c);
    Pack<VEC> sv = ldv<VEC>(s + p * s_stride + c);
    Pack<VEC> tv = ldv<VEC>(t + p * t_stride + c);
    Pack<VEC> x, dx, dsv, dtv;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float ex, dg;
      if (FAST) {
        scale_fast(kind, clamp, 1.0f / clamp, sv.v[e], ex, dg);
      } else {
        float g;
        log_scale(kind, clamp, sv.v[e], g, dg);
        ex = expf(g);
      }
      if (!inverse) {            // y = ex*x + t
        x.v[e] = FAST ? __fdividef(y.v[e] - tv.v[e], ex) : (y.v[e] - tv.v[e]) / ex;
        dx.v[e] = dy.v[e] * ex;
        dsv.v[e] = dy.v[e] * x.v[e] * ex * dg;
        dtv.v[e] = dy.v[e];
      } else {                   // y = (x - t)/ex
        x.v[e] = y.v[e] * ex + tv.v[e];
        float q = FAST ? __fdividef(dy.v[e], ex) : dy.v[e] / ex;
        dx.v[e] = q;
        dsv.v[e] = -dy.v[e] * y.v[e] * dg;
        dtv.v[e] = -q;
      }
    }
    stv<VEC>(u + p * u_stride + c, x);
    stv<VEC>(du + p * du_stride + c, dx);
    stv<VEC>(ds + p * ds_stride + c, dsv);
    stv<VEC>(dt + p * dt_stride + c, dtv);
    if (xbf != nullptr) stv<VEC>(xbf + p * L + c, x);
  }
}

// ---- the last half-step of a coupling block fused with the channel permutation that follows it (value pass), and the
// permutation's undo fused with the backward of that half-step (backward pass).  FrEIA: GLOWCouplingBlock followed by
// PermuteRandom (/root/reference/archs.py:61-68).  Standalone, the half-step rewrites half of the trunk in place and the
// permutation then moves all of it (4T bytes per block, 7T with the gradient in the backward pass); fused, every channel is
// read once and written once to its permuted place (3T / 5T).  One thread = 4 consecutive OUTPUT channels of a pixel.
//   out[p][i] = f(in[p][map[i]])   where f is the half-step if map[i] lies in the active range [c0, c0 + L), else identity
template <bool FAST>
__global__ void __launch_bounds__(256) coupling_apply_permute_kernel(const float* __restrict__ in, float* __restrict__ out, long long npix, int C,
                                                                     const int32_t* __restrict__ map, int c0, int L,
                                                                     const float* __restrict__ s, int s_stride, const float* __restrict__ t, int t_stride,
                                                                     int kind, float clamp, int inverse, __nv_bfloat16* __restrict__ bf, int bc0, int bc1) {
  pdl_wait();
  pdl_trigger();
  const int Cv = C / 4;
  const long long total = npix * Cv;
  const float inv_clamp = 1.0f / clamp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Cv, p, c);
    c *= 4;
    const float* row = in + p * C;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int sc = __ldg(map + c + e);
      float x = row[sc];
      const int j = sc - c0;
      if (j >= 0 && j < L) {
        const float sv = s[p * s_stride + j], tv = t[p * t_stride + j];
        float ex, dg;
        if (FAST) {
          scale_fast(kind, clamp, inv_clamp, sv, ex, dg);
          x = inverse ? __fdividef(x - tv, ex) : fmaf(ex, x, tv);
        } else {
          float g;
          log_scale(kind, clamp, sv, g, dg);
          ex = expf(g);
          x = inverse ? (x - tv) / ex : ex * x + tv;
        }
      }
      v[e] = x;
    }
    const float4 o = make_float4(v[0], v[1], v[2], v[3]);
    store4(out + p * C + c, o);
    if (bf != nullptr && c >= bc0 && c < bc1) store4(bf + p * (long long)(bc1 - bc0) + (c - bc0), o);
  }
}

// y_in / dy_in: trunk and gradient in the PERMUTED layout (the block's output after the permutation); map undoes the
// permutation (out channel i <- in channel map[i]).  Output channels [c0, c0 + L) (c0, L multiples of 4) are the half-step's:
// coupling_bwd_kernel's arithmetic; the others are moved unchanged.
template <typename TO, bool FAST>
__global__ void __launch_bounds__(256) coupling_bwd_unpermute_kernel(const float* __restrict__ y_in, const float* __restrict__ dy_in,
                                                                     float* __restrict__ x_out, float* __restrict__ dx_out, long long npix, int C,
                                                                     const int32_t* __restrict__ map, int c0, int L,
                                                                     const float* __restrict__ s, int s_stride, const float* __restrict__ t, int t_stride,
                                                                     int kind, float clamp, int inverse, TO* __restrict__ ds, int ds_stride,
                                                                     TO* __restrict__ dt, int dt_stride) {
  pdl_wait();
  pdl_trigger();
  const int Cv = C / 4;
  const long long total = npix * Cv;
  const float inv_clamp = 1.0f / clamp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Cv, p, c);
    c *= 4;
    const int m0 = __ldg(map + c), m1 = __ldg(map + c + 1), m2 = __ldg(map + c + 2), m3 = __ldg(map + c + 3);
    const float* ry = y_in + p * C;
    const float* rd = dy_in + p * C;
    const float y[4] = {ry[m0], ry[m1], ry[m2], ry[m3]};
    const float dy[4] = {rd[m0], rd[m1], rd[m2], rd[m3]};
    const int j = c - c0;
    if (j >= 0 && j < L) {
      const Pack<4> sv = ldv<4>(s + p * s_stride + j), tv = ldv<4>(t + p * t_stride + j);
      Pack<4> x, dx, dsv, dtv;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float ex, dg;
        if (FAST) {
          scale_fast(kind, clamp, inv_clamp, sv.v[e], ex, dg);
        } else {
          float g;
          log_scale(kind, clamp, sv.v[e], g, dg);
          ex = expf(g);
        }
        if (!inverse) {            // y = ex*x + t
          x.v[e] = FAST ? __fdividef(y[e] - tv.v[e], ex) : (y[e] - tv.v[e]) / ex;
          dx.v[e] = dy[e] * ex;
          dsv.v[e] = dy[e] * x.v[e] * ex * dg;
          dtv.v[e] = dy[e];
        } else {                   // y = (x - t)/ex
          x.v[e] = y[e] * ex + tv.v[e];
          const float q = FAST ? __fdividef(dy[e], ex) : dy[e] / ex;
          dx.v[e] = q;
          dsv.v[e] = -dy[e] * y[e] * dg;
          dtv.v[e] = -q;
        }
      }
      stv<4>(x_out + p * C + c, x);
      stv<4>(dx_out + p * C + c, dx);
      stv<4>(ds + p * ds_stride + j, dsv);
      stv<4>(dt + p * dt_stride + j, dtv);
    } else {
      store4(x_out + p * C + c, make_float4(y[0], y[1], y[2], y[3]));
      store4(dx_out + p * C + c, make_float4(dy[0], dy[1], dy[2], dy[3]));
    }
  }
}

template <int VEC, typename TO>
__global__ void __launch_bounds__(256) cast_slice_kernel(const float* __restrict__ in, int in_stride, long long npix, int L,
                                                         float scale, TO* __restrict__ out, int out_stride) {
  pdl_wait();
  pdl_trigger();
  const int Lv = L / VEC;
  const long long total = npix * Lv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Lv, p, c);
    c *= VEC;
    Pack<VEC> v = ldv<VEC>(in + p * in_stride + c);
#pragma unroll
    for (int e = 0; e < VEC; ++e) v.v[e] *= scale;
    stv<VEC>(out + p * out_stride + c, v);
  }
}

// VEC = 4: four channels per thread (16 / 8-byte accesses; the launchers check alignment), VEC = 1: any slice
template <int VEC, typename TY, typename TO>
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* d, int d_stride, const TY* __restrict__ y, int y_stride,
                                                      TO* out, int out_stride, long long npix, int L, int act, float slope) {
  pdl_wait();
  pdl_trigger();
  const int Lv = L / VEC;
  const long long total = npix * Lv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Lv, p, c);
    c *= VEC;
    if (VEC == 4) {
      const float4 yv = load4(y + p * y_stride + c), dv = load4(d + p * d_stride + c);
      store4(out + p * out_stride + c, make_float4(dv.x * act_grad(act, slope, yv.x), dv.y * act_grad(act, slope, yv.y),
                                                   dv.z * act_grad(act, slope, yv.z), dv.w * act_grad(act, slope, yv.w)));
    } else {
      float g = act_grad(act, slope, to_f32(y[p * y_stride + c]));
      out[p * out_stride + c] = from_f32<TO>(d[p * d_stride + c] * g);
    }
  }
}

template <int VEC, typename TA>
__global__ void __launch_bounds__(256) axpy_slice_kernel(float* __restrict__ out, int out_stride, const TA* __restrict__ a,
                                                         int a_stride, long long npix, int L, float alpha) {
  pdl_wait();
  pdl_trigger();
  const int Lv = L / VEC;
  const long long total = npix * Lv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c;
    long long p;
    split_index(idx, Lv, p, c);
    c *= VEC;
    if (VEC == 4) {
      const float4 av = load4(a + p * a_stride + c);
      float4 o = load4(out + p * out_stride + c);
      o.x += alpha * av.x; o.y += alpha * av.y; o.z += alpha * av.z; o.w += alpha * av.w;
      store4(out + p * out_stride + c, o);
    } else {
      out[p * out_stride + c] += alpha * to_f32(a[p * a_stride + c]);
    }
  }
}

// ---- column sums (bias gradients): partial[chunk][N] then fixed-order finish
constexpr int CS_COLS = 64, CS_ROWS = 4;
template <typename T>
__global__ void __launch_bounds__(CS_COLS* CS_ROWS) colsum_partial_kernel(const T* __restrict__ in, int in_stride, long long npix,
                                                                          int N, long long rows_per_chunk, float* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[CS_ROWS][CS_COLS];
  const int col = blockIdx.y * CS_COLS + threadIdx.x;
  const long long r0 = blockIdx.x * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > npix) r1 = npix;
  float acc = 0.f;
  if (col < N)
    for (long long r = r0 + threadIdx.y; r < r1; r += CS_ROWS) acc += to_f32(in[r * in_stride + col]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float sum = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < CS_ROWS; ++k) sum += red[k][threadIdx.x];
    partial[(long long)blockIdx.x * N + col] = sum;
  }
}
// Vectorised variant: every thread owns one 16-byte column group (8 bf16 / 4 fp32) and walks the rows of its
// block's chunk with stride RPI (rows per iteration = 256 / groups); rows are read as full coalesced lines.
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_vec_kernel(const T* __restrict__ in, int in_stride, long long npix, int N,
                                                                 long long rows_per_chunk, float* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  constexpr int VEC = 16 / (int)sizeof(T);
  extern __shared__ float red[];                        // [RPI][G*VEC]
  const int G = N / VEC;
  const int RPI = 256 / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  const long long r0 = blockIdx.x * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > npix) r1 = npix;
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  if (rl < RPI) {
    long long r = r0 + rl;
    // four rows in flight per thread: the kernel is pure streaming, memory-level parallelism is what matters
    for (; r + 3 * RPI < r1; r += 4 * RPI) {
      float4 t[4][VEC / 4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const T* src = in + (r + (long long)u * RPI) * in_stride + g * VEC;
#pragma unroll
        for (int q = 0; q < VEC / 4; ++q) t[u][q] = load4(src + 4 * q);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < VEC / 4; ++q) {
          acc[4 * q] += t[u][q].x; acc[4 * q + 1] += t[u][q].y; acc[4 * q + 2] += t[u][q].z; acc[4 * q + 3] += t[u][q].w;
        }
    }
    for (; r < r1; r += RPI) {
      const T* src = in + r * in_stride + g * VEC;
#pragma unroll
      for (int q = 0; q < VEC / 4; ++q) {
        const float4 t = load4(src + 4 * q);
        acc[4 * q] += t.x; acc[4 * q + 1] += t.y; acc[4 * q + 2] += t.z; acc[4 * q + 3] += t.w;
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) red[rl * N + g * VEC + e] = acc[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += 256) {
    float sum = 0.f;
    for (int k = 0; k < RPI; ++k) sum += red[k * N + c];
    partial[(long long)blockIdx.x * N + c] = sum;
  }
}

// block = 32 columns x 8 chunk-lanes; fixed-order tree over the lanes => deterministic
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ partial, int chunks, int N,
                                                            float* __restrict__ out, int accumulate) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ky = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float sum = 0.f;
  if (col < N)
    for (int k = ky; k < chunks; k += 8) sum += partial[(long long)k * N + col];
  red[ky][cx] = sum;
  __syncthreads();
  if (ky == 0 && col < N) {
    float t = red[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][cx];
    out[col] = accumulate ? out[col] + t : t;
  }
}

static inline long long colsum_chunks(long long npix) {
  long long chunks = (npix + 255) / 256;          // >= 256 rows per chunk
  long long cap = (long long)sm_count() * 4;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  return chunks;
}

// ---- sum of squared differences with optional gradient (loss.reconstruction / latent_nll)
__global__ void __launch_bounds__(256) sqdiff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                             float gscale, float* __restrict__ grad, float* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - (b ? b[i] : 0.f);
    acc += d * d;
    if (grad) grad[i] = gscale * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    partial[blockIdx.x] = s;
  }
}
// ---- the forward half's whole loss in one pass over the network output y [B][C][HW] (lit_wrapper.py:45-48):
//   w_rec * mean((y[:, :L] - lr)^2)  +  w_nll * mean(y[:, L:]^2)      and its gradient w.r.t. y (full tensor)
__global__ void __launch_bounds__(256) inn_fwd_loss_partial_kernel(const float* __restrict__ y, const float* __restrict__ lr, int C, int L,
                                                                   long long HW, long long total, float g_rec, float g_nll,
                                                                   float* __restrict__ grad, float* __restrict__ partial, int nblk) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[2][8];
  float a_rec = 0.f, a_nll = 0.f;
  const long long chw = (long long)C * HW, lhw = (long long)L * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long b, r;
    if (chw <= 0x7fffffffLL) { int r32; split_index(i, (int)chw, b, r32); r = r32; }
    else { b = i / chw; r = i - b * chw; }
    const float v = y[i];
    float g;
    if (r < lhw) {
      const float d = v - lr[b * lhw + r];
      a_rec += d * d;
      g = g_rec * d;
    } else {
      a_nll += v * v;
      g = g_nll * v;
    }
    grad[i] = g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a_rec += __shfl_xor_sync(0xffffffffu, a_rec, o);
    a_nll += __shfl_xor_sync(0xffffffffu, a_nll, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a_rec; red[1][threadIdx.x >> 5] = a_nll; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < 8; ++k) { s0 += red[0][k]; s1 += red[1][k]; }
    partial[blockIdx.x] = s0;
    partial[nblk + blockIdx.x] = s1;
  }
}
// Sum of n block partials in double by ONE block of 256 threads: thread t adds partials t, t + 256, ... in order, then a
// fixed-order tree through shared memory (bit-reproducible; a single thread walking thousands of partials cost 25-80 us
// on the critical path of a step).
__device__ __forceinline__ double block_sum_256(const float* __restrict__ partial, int n, double (*red)[8]) {
  double a = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) a += (double)partial[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  __syncthreads();                                   // (red may still be read from a previous call)
  if ((threadIdx.x & 31) == 0) (*red)[threadIdx.x >> 5] = a;
  __syncthreads();
  double s = 0.0;
  for (int k = 0; k < 8; ++k) s += (*red)[k];
  return s;
}
__global__ void __launch_bounds__(256) inn_fwd_loss_finish_kernel(const float* __restrict__ partial, int n, float s_rec, float s_nll,
                                                                  float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ double red[8];
  const double a = block_sum_256(partial, n, &red);
  const double b = block_sum_256(partial + n, n, &red);
  if (threadIdx.x == 0) out[0] = (float)(a * (double)s_rec + b * (double)s_nll);
}

__global__ void __launch_bounds__(256) sqdiff_finish_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ double red[8];
  const double s = block_sum_256(partial, n, &red);
  if (threadIdx.x == 0) out[0] = (float)(s * (double)scale);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float wd, float bc1, float bc2_sqrt, float gscale) {
  pdl_wait();
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    float gr = g[i] * gscale + wd * w;             // torch.optim.Adam: L2 weight decay folded into the gradient
    float mi = b1 * m[i] + (1.f - b1) * gr;
    float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = w - (lr / bc1) * (mi / denom);
  }
}

// Graph-replayable Adam: the step count lives on the device.  state = {int step, float bc1, float sqrt(bc2)}.
__global__ void adam_tick_kernel(int* __restrict__ state, float b1, float b2) {
  pdl_wait();
  pdl_trigger();
  const int step = state[0] + 1;
  state[0] = step;
  float* f = reinterpret_cast<float*>(state);
  f[1] = 1.f - powf(b1, (float)step);
  f[2] = sqrtf(1.f - powf(b2, (float)step));
}

__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                       float wd, const int* __restrict__ state, float gscale,
                                                       const float* __restrict__ g2) {
  pdl_wait();
  pdl_trigger();
  const float bc1 = reinterpret_cast<const float*>(state)[1], bc2_sqrt = reinterpret_cast<const float*>(state)[2];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    const float gsum = g2 != nullptr ? g[i] + g2[i] : g[i];        // second gradient arena of the two-stream step
    float gr = gsum * gscale + wd * w;
    float mi = b1 * m[i] + (1.f - b1) * gr;
    float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = w - (lr / bc1) * (mi / denom);
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = (long long)sm_count() * 32;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace sininn

using namespace sininn;

extern "C" {

int sininn_coupling_apply(float* u, int u_stride, const float* s, int s_stride, const float* t, int t_stride,
                          long long npix, int L, int kind, float clamp, int inverse, void* u_bf16, int fast_math,
                          sininn_stream_t stream) {
  SININN_CHECK_ARG(u && s && t && npix > 0 && L > 0, "coupling_apply: bad arguments");
  SININN_CHECK_ARG(kind == SININN_GLOW || kind == SININN_IRN, "coupling_apply: unknown kind %d", kind);
  SININN_CHECK_ARG(clamp > 0.f, "coupling_apply: clamp must be positive");
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(u_bf16);
  bool v4 = (L % 4) == 0 && (u_stride % 4) == 0 && (s_stride % 4) == 0 && (t_stride % 4) == 0 && aligned16(u) &&
            aligned16(s) && aligned16(t) && (!bf || aligned8(bf));
  const long long total = npix * (v4 ? L / 4 : L);
  const int block = 256, grid = grid_for(total, block);
#define LAUNCH(V, F) launch_k(coupling_apply_kernel<V, F>, dim3(grid), dim3(block), 0, as_stream(stream), u, u_stride, s, s_stride, t, t_stride, npix, L, kind, clamp, inverse, bf)
  if (fast_math) { if (v4) LAUNCH(4, true); else LAUNCH(1, true); }
  else           { if (v4) LAUNCH(4, false); else LAUNCH(1, false); }
#undef LAUNCH
  SININN_CHECK_LAUNCH("coupling_apply");
  return SININN_OK;
}

int sininn_coupling_bwd(float* u, int u_stride, float* du, int du_stride, const float* s, int s_stride, const float* t,
                        int t_stride, long long npix, int L, int kind, float clamp, int inverse, void* ds_out,
                        int ds_stride, void* dt_out, int dt_stride, int out_dtype, void* x_bf16, int fast_math,
                        sininn_stream_t stream) {
  SININN_CHECK_ARG(u && du && s && t && ds_out && dt_out && npix > 0 && L > 0, "coupling_bwd: bad arguments");
  SININN_CHECK_ARG(kind == SININN_GLOW || kind == SININN_IRN, "coupling_bwd: unknown kind %d", kind);
  SININN_CHECK_ARG(out_dtype == SININN_F32 || out_dtype == SININN_BF16, "coupling_bwd: bad out_dtype");
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(x_bf16);
  const bool f32 = out_dtype == SININN_F32;
  bool v4 = (L % 4) == 0 && (u_stride % 4) == 0 && (du_stride % 4) == 0 && (s_stride % 4) == 0 && (t_stride % 4) == 0 &&
            (ds_stride % 4) == 0 && (dt_stride % 4) == 0 && aligned16(u) && aligned16(du) && aligned16(s) && aligned16(t) &&
            (f32 ? (aligned16(ds_out) && aligned16(dt_out)) : (aligned8(ds_out) && aligned8(dt_out))) && (!bf || aligned8(bf));
  const long long total = npix * (v4 ? L / 4 : L);
  const int block = 256, grid = grid_for(total, block);
  cudaStream_t st = as_stream(stream);
#define LAUNCH2(V, T, F)                                                                                                \
  launch_k(coupling_bwd_kernel<V, T, F>, dim3(grid), dim3(block), 0, st, u, u_stride, du, du_stride, s, s_stride, t, t_stride, npix, L, kind, \
                                                    clamp, inverse, reinterpret_cast<T*>(ds_out), ds_stride,            \
                                                    reinterpret_cast<T*>(dt_out), dt_stride, bf)
#define LAUNCH(V, T) do { if (fast_math) LAUNCH2(V, T, true); else LAUNCH2(V, T, false); } while (0)
  if (f32) { if (v4) LAUNCH(4, float); else LAUNCH(1, float); }
  else     { if (v4) LAUNCH(4, __nv_bfloat16); else LAUNCH(1, __nv_bfloat16); }
#undef LAUNCH
#undef LAUNCH2
  SININN_CHECK_LAUNCH("coupling_bwd");
  return SININN_OK;
}

int sininn_coupling_apply_permute(const float* in, float* out, long long npix, int C, const int32_t* chan_map, int c0, int L,
                                  const float* s, int s_stride, const float* t, int t_stride, int kind, float clamp, int inverse,
                                  void* bf16_out, int bc0, int bc1, int fast_math, sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && chan_map && s && t && npix > 0 && C > 0 && in != out, "coupling_apply_permute: bad arguments");
  SININN_CHECK_ARG(kind == SININN_GLOW || kind == SININN_IRN, "coupling_apply_permute: unknown kind %d", kind);
  SININN_CHECK_ARG(clamp > 0.f && c0 >= 0 && L > 0 && c0 + L <= C, "coupling_apply_permute: bad active range / clamp");
  SININN_CHECK_ARG((C % 4) == 0 && aligned16(out), "coupling_apply_permute: needs C %% 4 == 0 and a 16-byte aligned output");
  __nv_bfloat16* bf = reinterpret_cast<__nv_bfloat16*>(bf16_out);
  if (bf) SININN_CHECK_ARG(0 <= bc0 && bc0 < bc1 && bc1 <= C && (bc0 % 4) == 0 && (bc1 % 4) == 0 && aligned8(bf), "coupling_apply_permute: bad bf16 channel range");
  const long long total = npix * (C / 4);
  const int block = 256, grid = grid_for(total, block);
  if (fast_math) launch_k(coupling_apply_permute_kernel<true>, dim3(grid), dim3(block), 0, as_stream(stream), in, out, npix, C, chan_map, c0, L, s, s_stride, t, t_stride, kind, clamp, inverse, bf, bc0, bc1);
  else launch_k(coupling_apply_permute_kernel<false>, dim3(grid), dim3(block), 0, as_stream(stream), in, out, npix, C, chan_map, c0, L, s, s_stride, t, t_stride, kind, clamp, inverse, bf, bc0, bc1);
  SININN_CHECK_LAUNCH("coupling_apply_permute");
  return SININN_OK;
}

int sininn_coupling_bwd_unpermute(const float* y_in, const float* dy_in, float* x_out, float* dx_out, long long npix, int C,
                                  const int32_t* chan_map, int c0, int L, const float* s, int s_stride, const float* t, int t_stride,
                                  int kind, float clamp, int inverse, void* ds_out, int ds_stride, void* dt_out, int dt_stride,
                                  int out_dtype, int fast_math, sininn_stream_t stream) {
  SININN_CHECK_ARG(y_in && dy_in && x_out && dx_out && chan_map && s && t && ds_out && dt_out && npix > 0 && C > 0, "coupling_bwd_unpermute: bad arguments");
  SININN_CHECK_ARG(y_in != x_out && dy_in != dx_out, "coupling_bwd_unpermute: cannot run in place");
  SININN_CHECK_ARG(kind == SININN_GLOW || kind == SININN_IRN, "coupling_bwd_unpermute: unknown kind %d", kind);
  SININN_CHECK_ARG(out_dtype == SININN_F32 || out_dtype == SININN_BF16, "coupling_bwd_unpermute: bad out_dtype");
  const bool f32 = out_dtype == SININN_F32;
  SININN_CHECK_ARG(clamp > 0.f && c0 >= 0 && L > 0 && c0 + L <= C && (C % 4) == 0 && (c0 % 4) == 0 && (L % 4) == 0,
                   "coupling_bwd_unpermute: C, c0 and L must be multiples of 4 (C=%d c0=%d L=%d)", C, c0, L);
  SININN_CHECK_ARG(aligned16(x_out) && aligned16(dx_out) && aligned16(s) && aligned16(t) && (s_stride % 4) == 0 && (t_stride % 4) == 0 &&
                   (ds_stride % 4) == 0 && (dt_stride % 4) == 0 && (f32 ? (aligned16(ds_out) && aligned16(dt_out)) : (aligned8(ds_out) && aligned8(dt_out))),
                   "coupling_bwd_unpermute: misaligned operands");
  const long long total = npix * (C / 4);
  const int block = 256, grid = grid_for(total, block);
  cudaStream_t st = as_stream(stream);
#define LAUNCH(T, F)                                                                                                                     \
  launch_k(coupling_bwd_unpermute_kernel<T, F>, dim3(grid), dim3(block), 0, st, y_in, dy_in, x_out, dx_out, npix, C, chan_map, c0, L, s, s_stride, \
           t, t_stride, kind, clamp, inverse, reinterpret_cast<T*>(ds_out), ds_stride, reinterpret_cast<T*>(dt_out), dt_stride)
  if (f32) { if (fast_math) LAUNCH(float, true); else LAUNCH(float, false); }
  else     { if (fast_math) LAUNCH(__nv_bfloat16, true); else LAUNCH(__nv_bfloat16, false); }
#undef LAUNCH
  SININN_CHECK_LAUNCH("coupling_bwd_unpermute");
  return SININN_OK;
}

int sininn_cast_slice(const float* in, int in_stride, long long npix, int L, float scale, void* out, int out_dtype,
                      int out_stride, sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && npix > 0 && L > 0, "cast_slice: bad arguments");
  SININN_CHECK_ARG(out_dtype == SININN_F32 || out_dtype == SININN_BF16, "cast_slice: bad out_dtype");
  const bool f32 = out_dtype == SININN_F32;
  bool v4 = (L % 4) == 0 && (in_stride % 4) == 0 && (out_stride % 4) == 0 && aligned16(in) && (f32 ? aligned16(out) : aligned8(out));
  const long long total = npix * (v4 ? L / 4 : L);
  const int block = 256, grid = grid_for(total, block);
  cudaStream_t st = as_stream(stream);
  if (f32) {
    if (v4) launch_k(cast_slice_kernel<4, float>, dim3(grid), dim3(block), 0, st, in, in_stride, npix, L, scale, (float*)out, out_stride);
    else launch_k(cast_slice_kernel<1, float>, dim3(grid), dim3(block), 0, st, in, in_stride, npix, L, scale, (float*)out, out_stride);
  } else {
    if (v4) launch_k(cast_slice_kernel<4, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, in, in_stride, npix, L, scale, (__nv_bfloat16*)out, out_stride);
    else launch_k(cast_slice_kernel<1, __nv_bfloat16>, dim3(grid), dim3(block), 0, st, in, in_stride, npix, L, scale, (__nv_bfloat16*)out, out_stride);
  }
  SININN_CHECK_LAUNCH("cast_slice");
  return SININN_OK;
}

int sininn_act_bwd(const float* d, int d_stride, const void* y, int y_dtype, int y_stride, void* out, int out_dtype,
                   int out_stride, long long npix, int L, int act, float slope, sininn_stream_t stream) {
  SININN_CHECK_ARG(d && y && out && npix > 0 && L > 0, "act_bwd: bad arguments");
  SININN_CHECK_ARG((y_dtype == SININN_F32 || y_dtype == SININN_BF16) && (out_dtype == SININN_F32 || out_dtype == SININN_BF16),
                   "act_bwd: bad dtype");
  const int esy = y_dtype == SININN_F32 ? 4 : 2, eso = out_dtype == SININN_F32 ? 4 : 2;
  const bool v4 = (L % 4) == 0 && (d_stride % 4) == 0 && (y_stride % 4) == 0 && (out_stride % 4) == 0 && aligned16(d) &&
                  (reinterpret_cast<uintptr_t>(y) % (4 * esy)) == 0 && (reinterpret_cast<uintptr_t>(out) % (4 * eso)) == 0;
  const long long total = npix * (v4 ? L / 4 : L);
  const int block = 256, grid = grid_for(total, block);
  cudaStream_t st = as_stream(stream);
#define LAUNCH2(V, TY, TO) launch_k(act_bwd_kernel<V, TY, TO>, dim3(grid), dim3(block), 0, st, d, d_stride, (const TY*)y, y_stride, (TO*)out, out_stride, npix, L, act, slope)
#define LAUNCH(TY, TO) do { if (v4) LAUNCH2(4, TY, TO); else LAUNCH2(1, TY, TO); } while (0)
  if (y_dtype == SININN_F32) { if (out_dtype == SININN_F32) LAUNCH(float, float); else LAUNCH(float, __nv_bfloat16); }
  else                       { if (out_dtype == SININN_F32) LAUNCH(__nv_bfloat16, float); else LAUNCH(__nv_bfloat16, __nv_bfloat16); }
#undef LAUNCH
#undef LAUNCH2
  SININN_CHECK_LAUNCH("act_bwd");
  return SININN_OK;
}

int sininn_axpy_slice(float* out, int out_stride, const void* a, int a_dtype, int a_stride, long long npix, int L,
                      float alpha, sininn_stream_t stream) {
  SININN_CHECK_ARG(out && a && npix > 0 && L > 0, "axpy_slice: bad arguments");
  SININN_CHECK_ARG(a_dtype == SININN_F32 || a_dtype == SININN_BF16, "axpy_slice: bad dtype");
  const int esa = a_dtype == SININN_F32 ? 4 : 2;
  const bool v4 = (L % 4) == 0 && (out_stride % 4) == 0 && (a_stride % 4) == 0 && aligned16(out) && (reinterpret_cast<uintptr_t>(a) % (4 * esa)) == 0;
  const long long total = npix * (v4 ? L / 4 : L);
  const int block = 256, grid = grid_for(total, block);
#define LAUNCH(V, TA) launch_k(axpy_slice_kernel<V, TA>, dim3(grid), dim3(block), 0, as_stream(stream), out, out_stride, (const TA*)a, a_stride, npix, L, alpha)
  if (a_dtype == SININN_F32) { if (v4) LAUNCH(4, float); else LAUNCH(1, float); }
  else                       { if (v4) LAUNCH(4, __nv_bfloat16); else LAUNCH(1, __nv_bfloat16); }
#undef LAUNCH
  SININN_CHECK_LAUNCH("axpy_slice");
  return SININN_OK;
}

size_t sininn_colsum_workspace_bytes(long long npix, int N) {
  if (npix <= 0 || N <= 0) return 0;
  return (size_t)colsum_chunks(npix) * (size_t)N * sizeof(float);
}

int sininn_colsum(const void* in, int dtype, int in_stride, long long npix, int N, float* out, int accumulate,
                  void* workspace, size_t workspace_bytes, sininn_stream_t stream) {
  SININN_CHECK_ARG(in && out && npix > 0 && N > 0, "colsum: bad arguments");
  SININN_CHECK_ARG(dtype == SININN_F32 || dtype == SININN_BF16, "colsum: bad dtype");
  const long long chunks = colsum_chunks(npix);
  if (!workspace || workspace_bytes < (size_t)chunks * N * sizeof(float)) {
    set_error("colsum: workspace too small (%zu < %zu)", workspace_bytes, (size_t)chunks * N * sizeof(float));
    return SININN_EWORKSPACE;
  }
  const long long rows_per_chunk = (npix + chunks - 1) / chunks;
  dim3 grid((unsigned)chunks, (N + CS_COLS - 1) / CS_COLS), block(CS_COLS, CS_ROWS);
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  const int vec = dtype == SININN_F32 ? 4 : 8;
  const bool vec_ok = (N % vec) == 0 && (N / vec) <= 256 && (in_stride % vec) == 0 && aligned16(in);
  if (vec_ok) {
    const int rpi = 256 / (N / vec);
    const size_t sm = (size_t)rpi * N * sizeof(float);
    if (dtype == SININN_F32) launch_k(colsum_partial_vec_kernel<float>, dim3((unsigned)chunks), dim3(256), sm, st, (const float*)in, in_stride, npix, N, rows_per_chunk, partial);
    else launch_k(colsum_partial_vec_kernel<__nv_bfloat16>, dim3((unsigned)chunks), dim3(256), sm, st, (const __nv_bfloat16*)in, in_stride, npix, N, rows_per_chunk, partial);
  } else if (dtype == SININN_F32) launch_k(colsum_partial_kernel<float>, dim3(grid), dim3(block), 0, st, (const float*)in, in_stride, npix, N, rows_per_chunk, partial);
  else launch_k(colsum_partial_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, st, (const __nv_bfloat16*)in, in_stride, npix, N, rows_per_chunk, partial);
  launch_k(colsum_finish_kernel, dim3((N + 31) / 32), dim3(256), 0, st, partial, (int)chunks, N, out, accumulate);
  SININN_CHECK_LAUNCH("colsum");
  return SININN_OK;
}

size_t sininn_sqdiff_workspace_bytes(long long n) {
  (void)n;
  return (size_t)sm_count() * 8 * sizeof(float);
}

int sininn_sqdiff_nchw(const float* a, const float* b, long long n, float scale, float* loss_out, float* grad_out,
                       void* workspace, size_t workspace_bytes, sininn_stream_t stream) {
  SININN_CHECK_ARG(a && loss_out && n > 0, "sqdiff: bad arguments");
  int grid = grid_for(n, 256);
  int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  if (!workspace || workspace_bytes < (size_t)grid * sizeof(float)) {
    set_error("sqdiff: workspace too small");
    return SININN_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  launch_k(sqdiff_partial_kernel, dim3(grid), dim3(256), 0, st, a, b, n, 2.f * scale, grad_out, (float*)workspace);
  launch_k(sqdiff_finish_kernel, dim3(1), dim3(256), 0, st, (const float*)workspace, grid, scale, loss_out);
  SININN_CHECK_LAUNCH("sqdiff");
  return SININN_OK;
}

int sininn_inn_fwd_loss(const float* y, const float* lr, int B, int C, int L, long long HW, float w_rec, float w_nll,
                        float* loss_out, float* grad_out, void* workspace, size_t workspace_bytes, sininn_stream_t stream) {
  SININN_CHECK_ARG(y && lr && loss_out && grad_out && B > 0 && C > 0 && L > 0 && L <= C && HW > 0, "inn_fwd_loss: bad arguments");
  const long long total = (long long)B * C * HW;
  int grid = grid_for(total, 256);
  int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  if (!workspace || workspace_bytes < (size_t)2 * grid * sizeof(float)) {
    set_error("inn_fwd_loss: workspace too small");
    return SININN_EWORKSPACE;
  }
  const float s_rec = w_rec / (float)((double)B * L * HW);
  const float s_nll = (C > L) ? w_nll / (float)((double)B * (C - L) * HW) : 0.f;
  cudaStream_t st = as_stream(stream);
  launch_k(inn_fwd_loss_partial_kernel, dim3(grid), dim3(256), 0, st, y, lr, C, L, HW, total, 2.f * s_rec, 2.f * s_nll, grad_out,
           (float*)workspace, grid);
  launch_k(inn_fwd_loss_finish_kernel, dim3(1), dim3(256), 0, st, (const float*)workspace, grid, s_rec, s_nll, loss_out);
  SININN_CHECK_LAUNCH("inn_fwd_loss");
  return SININN_OK;
}

int sininn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                     sininn_stream_t stream) {
  SININN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  const int block = 256, grid = grid_for(n, block);
  launch_k(adam_kernel, dim3(grid), dim3(block), 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                     weight_decay, bc1, sqrtf(bc2), grad_scale);
  SININN_CHECK_LAUNCH("adam_step");
  return SININN_OK;
}

int sininn_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, int* step_state, float grad_scale,
                         sininn_stream_t stream) {
  SININN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_state && n > 0, "adam_step_dev: bad arguments");
  const int block = 256, grid = grid_for(n, block);
  launch_k(adam_tick_kernel, dim3(1), dim3(1), 0, as_stream(stream), step_state, beta1, beta2);
  launch_k(adam_dev_kernel, dim3(grid), dim3(block), 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                         weight_decay, step_state, grad_scale, (const float*)nullptr);
  SININN_CHECK_LAUNCH("adam_step_dev");
  return SININN_OK;
}

int sininn_adam_step_dev2(float* param, const float* grad, const float* grad_b, float* exp_avg, float* exp_avg_sq, long long n,
                          float lr, float beta1, float beta2, float eps, float weight_decay, int* step_state,
                          float grad_scale, sininn_stream_t stream) {
  SININN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step_state && n > 0, "adam_step_dev2: bad arguments");
  const int block = 256, grid = grid_for(n, block);
  launch_k(adam_tick_kernel, dim3(1), dim3(1), 0, as_stream(stream), step_state, beta1, beta2);
  launch_k(adam_dev_kernel, dim3(grid), dim3(block), 0, as_stream(stream), param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                         weight_decay, step_state, grad_scale, grad_b);
  SININN_CHECK_LAUNCH("adam_step_dev2");
  return SININN_OK;
}

}  // extern "C"
