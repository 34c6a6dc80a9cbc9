// Kernels at the boundary of the INN (SURVEY.md section 8f rows and the optional FrEIA operators):
//   * latent_to_nhwc     cat((lr, z), dim=1) of lit_wrapper.py:41-42 / 110-111 folded into the channels-last entry
//                        of the inverse pass; z is read, or drawn on the device (Philox4x32-10 + Box-Muller)
//   * mmd                loss.mmd (loss.py:9-36), device-agnostic, value and gradient w.r.t. x
//   * channel_affine     FrEIA ActNorm (offered, commented out, at archs.py:40-44): y = x * exp(scale_c) + bias_c
//   * logscale_sum       log-determinant of a coupling half: sum over a sample of the clamped log-scales
#include "common.cuh"

namespace sininn {

static inline int grid_cap(long long total, int block, int per_sm) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)sm_count() * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------- Philox4x32-10 (counter-based, one call = 4 words)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// standard normal for element `idx` of stream (seed, offset)
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long idx) {
  uint32_t r[4];
  philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float u1 = ((float)r[0] + 1.0f) * 2.3283064365386963e-10f;        // (0, 1]
  const float u2 = (float)r[1] * 2.3283064365386963e-10f;
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// out[b][p][i] = src[b][map[i]][p] where src = cat(lr [B][L][HW], z [B][Z][HW]); z == nullptr: z = temp * N(0,1) drawn here
__global__ void __launch_bounds__(256) latent_to_nhwc_kernel(const float* __restrict__ lr, int L, const float* __restrict__ z, int Z, int HW,
                                                             const int32_t* __restrict__ map, float* __restrict__ out,
                                                             __nv_bfloat16* __restrict__ bf, int bc0, int bc1,
                                                             unsigned long long seed, unsigned long long offset, float temp,
                                                             float* __restrict__ z_out, const int32_t* __restrict__ step_ptr,
                                                             unsigned long long step_stride) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  if (step_ptr != nullptr) offset += (unsigned long long)(uint32_t)__ldg(step_ptr) * step_stride;   // a fresh z every replayed step
  const int C = L + Z;
  const long long b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = c0 + ty + 8 * r, p = p0 + tx;
    if (c < C && p < HW) {
      const int cs = map ? map[c] : c;
      float v;
      if (cs < L) {
        v = __ldcs(lr + (b * L + cs) * (long long)HW + p);
      } else {
        const long long zi = (b * Z + (cs - L)) * (long long)HW + p;
        if (z != nullptr) {
          v = __ldcs(z + zi);
        } else {
          v = temp * philox_normal(seed, offset + (unsigned long long)zi);
          if (z_out != nullptr) z_out[zi] = v;
        }
      }
      tile[ty + 8 * r][tx] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + 8 * r, c = c0 + tx;
    if (c < C && p < HW) {
      const float v = tile[tx][ty + 8 * r];
      out[(b * HW + p) * (long long)C + c] = v;
      if (bf != nullptr && c >= bc0 && c < bc1) bf[(b * HW + p) * (long long)(bc1 - bc0) + (c - bc0)] = __float2bfloat16_rn(v);
    }
  }
}

// ---------------------------------------------------------------- loss.mmd
// Inverse multiquadric kernels of loss.py:10-13: (C, a) pairs; term(d) = C^a * ((C + d) / a)^-a
__device__ __forceinline__ void mmd_kernels(int rev, float (&Cs)[3], float (&As)[3]) {
  if (rev) { Cs[0] = 0.2f; Cs[1] = 0.2f; Cs[2] = 0.2f; As[0] = 0.1f; As[1] = 0.5f; As[2] = 2.f; }
  else     { Cs[0] = 0.2f; Cs[1] = 1.5f; Cs[2] = 3.0f; As[0] = 2.f; As[1] = 2.f; As[2] = 2.f; }
}

constexpr int MMD_CHUNK = 64;
// partial[blk][which][i][j] = sum over this block's columns of P[i][d] * Q[j][d]; which 0: x.x, 1: y.y, 2: x.y
__global__ void __launch_bounds__(256) mmd_gram_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, int b, long long D,
                                                               float* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  __shared__ float xs[64][MMD_CHUNK + 1], ys[64][MMD_CHUNK + 1];
  const int nout = 3 * b * b;
  float acc[48];                                    // 3 * 64 * 64 / 256
#pragma unroll
  for (int k = 0; k < 48; ++k) acc[k] = 0.f;
  const long long nchunks = (D + MMD_CHUNK - 1) / MMD_CHUNK;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long d0 = ch * MMD_CHUNK;
    for (int e = threadIdx.x; e < b * MMD_CHUNK; e += 256) {
      const int i = e / MMD_CHUNK, d = e % MMD_CHUNK;
      const bool ok = d0 + d < D;
      xs[i][d] = ok ? __ldg(x + (long long)i * D + d0 + d) : 0.f;
      ys[i][d] = ok ? __ldg(y + (long long)i * D + d0 + d) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 48; ++k) {
      const int o = threadIdx.x + 256 * k;
      if (o < nout) {
        const int which = o / (b * b), i = (o % (b * b)) / b, j = o % b;
        const float* pa = which == 1 ? ys[i] : xs[i];
        const float* pb = which == 0 ? xs[j] : ys[j];
        float s = 0.f;
#pragma unroll 16
        for (int d = 0; d < MMD_CHUNK; ++d) s += pa[d] * pb[d];
        acc[k] += s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 48; ++k) {
    const int o = threadIdx.x + 256 * k;
    if (o < nout) partial[(long long)blockIdx.x * nout + o] = acc[k];
  }
}

__global__ void __launch_bounds__(256) mmd_gram_sum_kernel(const float* __restrict__ partial, int nblk, int nout, float* __restrict__ gram) {
  pdl_wait();
  pdl_trigger();
  const int o = blockIdx.x * 256 + threadIdx.x;
  if (o >= nout) return;
  float s = 0.f;
  for (int k = 0; k < nblk; ++k) s += partial[(long long)k * nout + o];       // fixed order
  gram[o] = s;
}

// one block: loss and the two coefficient matrices of the gradient: A = dL/d dxx + its transpose, Bm = dL/d dxy
__global__ void __launch_bounds__(256) mmd_finish_kernel(const float* __restrict__ gram, int b, int rev, float scale, float* __restrict__ loss,
                                                         float* __restrict__ A, float* __restrict__ Bm) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[256];
  __shared__ float S[64 * 64], Gxy[64 * 64];
  float Cs[3], As[3];
  mmd_kernels(rev, Cs, As);
  const float* xx = gram; const float* yy = gram + b * b; const float* xy = gram + 2 * b * b;
  const float inv = scale / (float)(b * b);
  float local = 0.f;
  for (int e = threadIdx.x; e < b * b; e += 256) {
    const int i = e / b, j = e % b;
    const float rdxx = xx[i * b + i] + xx[j * b + j] - 2.f * xx[e];
    const float rdyy = yy[i * b + i] + yy[j * b + j] - 2.f * yy[e];
    const float rdxy = xx[i * b + i] + yy[j * b + j] - 2.f * xy[e];
    const float dxx = fmaxf(rdxx, 0.f), dyy = fmaxf(rdyy, 0.f), dxy = fmaxf(rdxy, 0.f);
    float XX = 0.f, YY = 0.f, XY = 0.f, gxx = 0.f, gxy = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float ca = powf(Cs[k], As[k]);
      XX += ca * powf((Cs[k] + dxx) / As[k], -As[k]);
      YY += ca * powf((Cs[k] + dyy) / As[k], -As[k]);
      XY += ca * powf((Cs[k] + dxy) / As[k], -As[k]);
      gxx -= ca * powf((Cs[k] + dxx) / As[k], -As[k] - 1.f);     // d term / d d
      gxy -= ca * powf((Cs[k] + dxy) / As[k], -As[k] - 1.f);
    }
    local += XX + YY - 2.f * XY;
    S[e] = rdxx >= 0.f ? inv * gxx : 0.f;                        // torch.clamp passes the gradient on [min, max]
    Gxy[e] = rdxy >= 0.f ? -2.f * inv * gxy : 0.f;               // (the loss holds -2 XY)
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = inv * red[0];
  if (A == nullptr) return;
  // S <- Gxx + Gxx^T (x_i enters dxx_ij and dxx_ji)
  for (int e = threadIdx.x; e < b * b; e += 256) {
    const int i = e / b, j = e % b;
    if (i < j) { const float t = S[e] + S[j * b + i]; S[e] = t; S[j * b + i] = t; }
    else if (i == j) S[e] = 2.f * S[e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < b * b; e += 256) { A[e] = S[e]; Bm[e] = Gxy[e]; }
}

template <int BMAX>
__global__ void __launch_bounds__(256) mmd_grad_kernel(const float* __restrict__ x, const float* __restrict__ y, int b, long long D,
                                                       const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ grad) {
  pdl_wait();
  pdl_trigger();
  __shared__ float As[BMAX * BMAX], Bs[BMAX * BMAX];
  for (int e = threadIdx.x; e < b * b; e += 256) { As[e] = A[e]; Bs[e] = Bm[e]; }
  __syncthreads();
  for (long long d = blockIdx.x * 256LL + threadIdx.x; d < D; d += (long long)gridDim.x * 256) {
    float xr[BMAX], yr[BMAX];
#pragma unroll
    for (int j = 0; j < BMAX; ++j) {
      xr[j] = j < b ? __ldg(x + (long long)j * D + d) : 0.f;
      yr[j] = j < b ? __ldg(y + (long long)j * D + d) : 0.f;
    }
    // grad_i = 2 * sum_j [ S_ij (x_i - x_j) + Gxy_ij (x_i - y_j) ]: differences first (the expanded form
    // diag(rowsum) x - S x cancels catastrophically in fp32)
    for (int i = 0; i < b; ++i) {
      const float xi = __ldg(x + (long long)i * D + d);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < BMAX; ++j)
        if (j < b) s += As[i * b + j] * (xi - xr[j]) + Bs[i * b + j] * (xi - yr[j]);
      grad[(long long)i * D + d] = 2.f * s;
    }
  }
}

// ---------------------------------------------------------------- per-channel affine (ActNorm)
__global__ void __launch_bounds__(256) channel_affine_kernel(float* __restrict__ u, long long npix, int C, const float* __restrict__ ls,
                                                             const float* __restrict__ bias, int inverse) {
  pdl_wait();
  pdl_trigger();
  const long long total = npix * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C);
    const float v = u[i];
    u[i] = inverse ? (v - __ldg(bias + c)) * expf(-__ldg(ls + c)) : v * expf(__ldg(ls + c)) + __ldg(bias + c);
  }
}

// Backward from the OUTPUT: u holds y -> x, du holds dy -> dx; per-block partial sums of d/dscale and d/dbias.
// Thread t owns channel t % C of the rows it visits, so its two sums stay in registers; partial[blk][2][256].
__global__ void __launch_bounds__(256) channel_affine_bwd_kernel(float* __restrict__ u, float* __restrict__ du, long long npix, int C,
                                                                 const float* __restrict__ ls, const float* __restrict__ bias, int inverse,
                                                                 float* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  const int rows_per_pass = 256 / C;                 // host guarantees C <= 256
  const int c = threadIdx.x % C, r = threadIdx.x / C;
  const bool active = r < rows_per_pass;
  const float s = expf(__ldg(ls + c)), bi = __ldg(bias + c);
  float ds = 0.f, db = 0.f;
  if (active) {
    for (long long p = (long long)blockIdx.x * rows_per_pass + r; p < npix; p += (long long)gridDim.x * rows_per_pass) {
      const long long i = p * C + c;
      const float y = u[i], dy = du[i];
      if (!inverse) {            // y = x * s + b
        u[i] = (y - bi) / s;
        du[i] = dy * s;
        ds += dy * (y - bi);
        db += dy;
      } else {                   // y = (x - b) / s
        u[i] = y * s + bi;
        du[i] = dy / s;
        ds -= dy * y;
        db -= dy / s;
      }
    }
  }
  __shared__ float sd[256], sb[256];
  sd[threadIdx.x] = active ? ds : 0.f;
  sb[threadIdx.x] = active ? db : 0.f;
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f, bsum = 0.f;
    for (int k = 0; k < rows_per_pass; ++k) { a += sd[k * C + threadIdx.x]; bsum += sb[k * C + threadIdx.x]; }
    partial[(long long)blockIdx.x * 512 + threadIdx.x] = a;
    partial[(long long)blockIdx.x * 512 + 256 + threadIdx.x] = bsum;
  }
}

__global__ void __launch_bounds__(256) channel_affine_bwd_finish_kernel(const float* __restrict__ partial, int nblk, int C,
                                                                        float* __restrict__ dls, float* __restrict__ dbias, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const int c = threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < nblk; ++k) { a += partial[(long long)k * 512 + c]; b += partial[(long long)k * 512 + 256 + c]; }
  dls[c] = accumulate ? dls[c] + a : a;
  dbias[c] = accumulate ? dbias[c] + b : b;
}

// out[sample] (+)= sign * sum over the sample's pixels and L channels of g(s); one block per sample, fixed order
__global__ void __launch_bounds__(256) logscale_sum_kernel(const float* __restrict__ s, int s_stride, long long pix_per_sample, int L,
                                                           int kind, float clamp, float sign, float* __restrict__ out, int accumulate) {
  pdl_wait();
  pdl_trigger();
  const long long base = (long long)blockIdx.x * pix_per_sample;
  const long long total = pix_per_sample * L;
  float acc = 0.f;
  for (long long e = threadIdx.x; e < total; e += 256) {
    const long long p = e / L;
    const int c = (int)(e % L);
    float g, dg;
    log_scale(kind, clamp, s[(base + p) * s_stride + c], g, dg);
    acc += g;
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (accumulate ? out[blockIdx.x] : 0.f) + sign * red[0];
}

}  // namespace sininn

using namespace sininn;

extern "C" {

int sininn_latent_to_nhwc(const float* lr, int L, const float* z, int Z, int B, int HW, const int32_t* chan_map, float* out,
                          void* bf16_out, int c0, int c1, unsigned long long seed, unsigned long long offset, float temp,
                          float* z_out, const int32_t* step_ptr, unsigned long long step_stride, sininn_stream_t stream) {
  SININN_CHECK_ARG(lr && out && L > 0 && Z >= 0 && B > 0 && HW > 0, "latent_to_nhwc: bad arguments");
  SININN_CHECK_ARG(B <= 65535, "latent_to_nhwc: batch too large for grid.z");
  const int C = L + Z;
  if (bf16_out) SININN_CHECK_ARG(0 <= c0 && c0 < c1 && c1 <= C, "latent_to_nhwc: bad bf16 channel range");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  launch_k(latent_to_nhwc_kernel, grid, block, 0, as_stream(stream), lr, L, z, Z, HW, chan_map, out,
           reinterpret_cast<__nv_bfloat16*>(bf16_out), c0, c1, seed, offset, temp, z_out, step_ptr, step_stride);
  SININN_CHECK_LAUNCH("latent_to_nhwc");
  return SININN_OK;
}

size_t sininn_mmd_workspace_bytes(int b, long long D) {
  if (b <= 0 || b > 64 || D <= 0) return 0;
  const size_t nout = (size_t)3 * b * b;
  const size_t nblk = (size_t)sm_count() * 2;
  return (nblk * nout + nout + 2 * (size_t)b * b) * sizeof(float);
}

int sininn_mmd(const float* x, const float* y, int b, long long D, int rev, float scale, float* loss_out, float* grad_x_out,
               void* workspace, size_t workspace_bytes, sininn_stream_t stream) {
  SININN_CHECK_ARG(x && y && loss_out && D > 0, "mmd: bad arguments");
  SININN_CHECK_ARG(b >= 1 && b <= 64, "mmd: batch must be 1..64 (got %d)", b);
  const size_t need = sininn_mmd_workspace_bytes(b, D);
  if (!workspace || workspace_bytes < need) {
    set_error("mmd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return SININN_EWORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int nout = 3 * b * b;
  long long nchunks = (D + MMD_CHUNK - 1) / MMD_CHUNK;
  int nblk = sm_count() * 2;
  if (nblk > nchunks) nblk = (int)nchunks;
  float* partial = reinterpret_cast<float*>(workspace);
  float* gram = partial + (size_t)sm_count() * 2 * nout;
  float* A = gram + nout;
  float* Bm = A + b * b;
  launch_k(mmd_gram_partial_kernel, dim3(nblk), dim3(256), 0, st, x, y, b, D, partial);
  launch_k(mmd_gram_sum_kernel, dim3((nout + 255) / 256), dim3(256), 0, st, (const float*)partial, nblk, nout, gram);
  launch_k(mmd_finish_kernel, dim3(1), dim3(256), 0, st, (const float*)gram, b, rev, scale, loss_out,
           grad_x_out ? A : (float*)nullptr, grad_x_out ? Bm : (float*)nullptr);
  if (grad_x_out) {
    const int grid = grid_cap(D, 256, 8);
    if (b <= 8) launch_k(mmd_grad_kernel<8>, dim3(grid), dim3(256), 0, st, x, y, b, D, (const float*)A, (const float*)Bm, grad_x_out);
    else if (b <= 16) launch_k(mmd_grad_kernel<16>, dim3(grid), dim3(256), 0, st, x, y, b, D, (const float*)A, (const float*)Bm, grad_x_out);
    else if (b <= 32) launch_k(mmd_grad_kernel<32>, dim3(grid), dim3(256), 0, st, x, y, b, D, (const float*)A, (const float*)Bm, grad_x_out);
    else launch_k(mmd_grad_kernel<64>, dim3(grid), dim3(256), 0, st, x, y, b, D, (const float*)A, (const float*)Bm, grad_x_out);
  }
  SININN_CHECK_LAUNCH("mmd");
  return SININN_OK;
}

int sininn_channel_affine(float* u, long long npix, int C, const float* log_scale, const float* bias, int inverse,
                          sininn_stream_t stream) {
  SININN_CHECK_ARG(u && log_scale && bias && npix > 0 && C > 0, "channel_affine: bad arguments");
  launch_k(channel_affine_kernel, dim3(grid_cap(npix * C, 256, 16)), dim3(256), 0, as_stream(stream), u, npix, C, log_scale, bias, inverse);
  SININN_CHECK_LAUNCH("channel_affine");
  return SININN_OK;
}

size_t sininn_channel_affine_bwd_workspace_bytes(void) { return (size_t)sm_count() * 4 * 512 * sizeof(float); }

int sininn_channel_affine_bwd(float* u, float* du, long long npix, int C, const float* log_scale, const float* bias, int inverse,
                              float* dlog_scale, float* dbias, int accumulate, void* workspace, size_t workspace_bytes,
                              sininn_stream_t stream) {
  SININN_CHECK_ARG(u && du && log_scale && bias && dlog_scale && dbias && npix > 0, "channel_affine_bwd: bad arguments");
  SININN_CHECK_ARG(C >= 1 && C <= 256, "channel_affine_bwd: 1..256 channels supported (got %d)", C);
  if (!workspace || workspace_bytes < sininn_channel_affine_bwd_workspace_bytes()) {
    set_error("channel_affine_bwd: workspace too small");
    return SININN_EWORKSPACE;
  }
  const int rows = 256 / C;
  int nblk = sm_count() * 4;
  if ((long long)nblk * rows > npix) nblk = (int)((npix + rows - 1) / rows);
  cudaStream_t st = as_stream(stream);
  launch_k(channel_affine_bwd_kernel, dim3(nblk), dim3(256), 0, st, u, du, npix, C, log_scale, bias, inverse, (float*)workspace);
  launch_k(channel_affine_bwd_finish_kernel, dim3(1), dim3(256), 0, st, (const float*)workspace, nblk, C, dlog_scale, dbias, accumulate);
  SININN_CHECK_LAUNCH("channel_affine_bwd");
  return SININN_OK;
}

int sininn_logscale_sum(const float* s, int s_stride, int B, long long pix_per_sample, int L, int kind, float clamp, float sign,
                        float* out, int accumulate, sininn_stream_t stream) {
  SININN_CHECK_ARG(s && out && B > 0 && pix_per_sample > 0 && L > 0 && s_stride >= L, "logscale_sum: bad arguments");
  SININN_CHECK_ARG(kind == SININN_GLOW || kind == SININN_IRN, "logscale_sum: unknown coupling kind %d", kind);
  launch_k(logscale_sum_kernel, dim3(B), dim3(256), 0, as_stream(stream), s, s_stride, pix_per_sample, L, kind, clamp, sign, out, accumulate);
  SININN_CHECK_LAUNCH("logscale_sum");
  return SININN_OK;
}

}  // extern "C"
