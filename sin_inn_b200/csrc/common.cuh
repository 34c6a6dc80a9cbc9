// Shared helpers for libsininn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sininn.h"

namespace sininn {

void set_error(const char* fmt, ...);

#define SININN_CHECK_ARG(cond, ...)                         \
  do {                                                      \
    if (!(cond)) {                                          \
      ::sininn::set_error(__VA_ARGS__);                     \
      return SININN_EINVAL;                                 \
    }                                                       \
  } while (0)

#define SININN_CHECK_LAUNCH(name)                                               \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      ::sininn::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return SININN_ECUDA;                                                      \
    }                                                                           \
  } while (0)

static inline cudaStream_t as_stream(sininn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

int sm_count();

// ---- programmatic dependent launch (PDL).  Every kernel of this library is launched with the
// programmaticStreamSerialization attribute and starts with pdl_wait(): a kernel's CTAs may be scheduled (and run
// their prologue: barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel in the stream is
// still draining, but touch no global memory before the previous kernel has completed and flushed.  That hides the
// launch latency and prologue of ~560 back-to-back launches per training step.  SININN_PDL=0 switches it off.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// first statement (before any global-memory access) of every kernel launched through launch_k
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the next kernel's CTAs be scheduled as SMs free up (they still block in their own pdl_wait)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements -> fp32 (pointer must be 16 B / 8 B aligned for float / bf16)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// ---- coupling log-scales (fp32, accurate libm variants: the fp32 path must match the reference to 1e-4)
// GLOW: g(s) = clamp*0.636*atan(s/clamp), g'(s) = 0.636/(1+(s/clamp)^2)      (FrEIA pre-v0.2, literal 0.636)
// IRN : g(h) = clamp*(2*sigmoid(h)-1),   g'(h) = 2*clamp*sig*(1-sig)         (archs.py:153)
__device__ __forceinline__ void log_scale(int kind, float clamp, float raw, float& g, float& dg) {
  if (kind == SININN_GLOW) {
    float r = raw / clamp;
    g = clamp * 0.636f * atanf(r);
    dg = 0.636f / (1.0f + r * r);
  } else {
    float sg = 1.0f / (1.0f + expf(-raw));
    g = clamp * (2.0f * sg - 1.0f);
    dg = 2.0f * clamp * sg * (1.0f - sg);
  }
}

// ---- fast forms for the bf16 path (standalone coupling kernels with fast_math = 1 and the coupling epilogues of the tcgen05
// kernels): a degree-7 polynomial in t^2 for atan (1.6e-7 absolute on [0, 1], reciprocal argument beyond 1), ex2.approx,
// approximate division -- errors at fp32 rounding level, far inside that path's tolerance.  The accurate libm forms above cost
// ~60 ALU instructions per element, which makes the "bandwidth-bound" coupling kernels ALU-bound (16 lanes per clock and
// scheduler): measured 15 / 21 us per launch against 8 / 14 us of HBM time at the headline shape.
__device__ __forceinline__ float fast_atan(float r) {
  const float a = fabsf(r);
  const bool inv = a > 1.0f;
  const float t = inv ? __fdividef(1.0f, a) : a;
  const float z = t * t;
  float p = -0.004668773151934147f;
  p = fmaf(p, z, 0.02416618913412094f);
  p = fmaf(p, z, -0.0593671016395092f);
  p = fmaf(p, z, 0.09906096756458282f);
  p = fmaf(p, z, -0.14016585052013397f);
  p = fmaf(p, z, 0.19969235360622406f);
  p = fmaf(p, z, -0.33331960439682007f);
  p = fmaf(p, z, 0.9999998807907104f);
  p *= t;
  p = inv ? 1.5707963267948966f - p : p;
  return copysignf(p, r);
}
// e = exp(g(s)), dg = g'(s) for the GLOW clamp (log_scale above)
__device__ __forceinline__ void glow_scale_fast(float clamp, float inv_clamp, float sv, float& ex, float& dg) {
  const float r = sv * inv_clamp;
  ex = __expf(clamp * 0.636f * fast_atan(r));
  dg = __fdividef(0.636f, fmaf(r, r, 1.0f));
}

// e = exp(g), dg = g' for either coupling kind
__device__ __forceinline__ void scale_fast(int kind, float clamp, float inv_clamp, float raw, float& ex, float& dg) {
  if (kind == SININN_GLOW) {
    glow_scale_fast(clamp, inv_clamp, raw, ex, dg);
  } else {
    const float sg = __fdividef(1.0f, 1.0f + __expf(-raw));
    ex = __expf(clamp * (2.0f * sg - 1.0f));
    dg = 2.0f * clamp * sg * (1.0f - sg);
  }
}

// idx = p * n + c (0 <= c < n): 32-bit arithmetic whenever the index space allows it -- a 64-bit division is a ~100-instruction
// subroutine, several times the rest of an elementwise kernel's body
__device__ __forceinline__ void split_index(long long idx, int n, long long& p, int& c) {
  if (idx <= 0xffffffffLL) {
    const unsigned q = (unsigned)idx / (unsigned)n;
    p = q;
    c = (int)((unsigned)idx - q * (unsigned)n);
  } else {
    p = idx / n;
    c = (int)(idx - p * n);
  }
}

__device__ __forceinline__ float act_fwd(int act, float slope, float v) {
  if (act == SININN_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == SININN_ACT_LRELU) return v > 0.f ? v : slope * v;
  return v;
}
// derivative from the activation OUTPUT y (both activations preserve sign)
__device__ __forceinline__ float act_grad(int act, float slope, float y) {
  if (act == SININN_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == SININN_ACT_LRELU) return y > 0.f ? 1.f : slope;
  return 1.f;
}

}  // namespace sininn
