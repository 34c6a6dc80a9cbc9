/*
 * sininn.h -- C ABI of libsininn.so: the sm_100a (B200) kernels behind the
 * invertible-network hot path of paramhanji/sin-inn.
 *
 * The reference has no native boundary: its hot path is Python that calls
 * ATen/cuDNN (archs.py) and the un-vendored FrEIA package.  Each entry point
 * below names the reference code it replaces (paths relative to the reference
 * repository root).  Host bindings: sin_inn_b200/_lib.py (ctypes);
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless said otherwise
 *   - nothing allocates, synchronises or takes ownership; kernels are launched on
 *     the stream passed in (a cudaStream_t, passed as void*)
 *   - return value 0 = ok, negative = SININN_E*; sininn_last_error() gives text
 *   - "NHWC" = channels-last activation matrix [B*H*W pixels][C], addressed with an
 *     explicit pixel stride (in elements) so channel slices of a wider tensor work
 *   - dtype codes: SININN_F32 / SININN_BF16
 */
#ifndef SININN_H
#define SININN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SININN_OK            0
#define SININN_EINVAL       -1
#define SININN_ECUDA        -2
#define SININN_EUNSUPPORTED -3
#define SININN_EWORKSPACE   -4

#define SININN_F32  0
#define SININN_BF16 1

/* coupling kinds */
#define SININN_GLOW 0   /* log-scale g(s) = clamp*0.636*atan(s/clamp)   (FrEIA GLOWCouplingBlock, call site archs.py:61-64) */
#define SININN_IRN  1   /* log-scale g(h) = clamp*(2*sigmoid(h)-1)      (InvBlockExp, archs.py:153,156) */

/* activations fused into conv epilogues */
#define SININN_ACT_NONE  0
#define SININN_ACT_RELU  1   /* nn.ReLU in subnet_conv / subnet_conv_1x1, archs.py:11-17 */
#define SININN_ACT_LRELU 2   /* LeakyReLU(0.2) in DenseBlock, archs.py:82,89-93 */

typedef void* sininn_stream_t;

int         sininn_version(void);
const char* sininn_last_error(void);
/* sm count / compute capability of the current device; any pointer may be NULL */
int         sininn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- resampling
 * Replaces FrEIA IRevNetDownsampling (call sites archs.py:28-31,35-38):
 *   out[b,(dy*2+dx)*C+c,i,j] = in[b,c,2i+dy,2j+dx]
 * and HaarDownsampling.forward (archs.py:183-199): the four +-1 2x2 patterns of
 * archs.py:167-176, band-major output channel k*C+c; `scale` multiplies the
 * result (0.25 = reference forward, 1 = reference reverse; other values give
 * the transposes needed by the backward pass).
 * C,H,W always describe the FULL-resolution side.  rev=0: [B,C,H,W] ->
 * [B,4C,H/2,W/2]; rev=1: the inverse mapping.  mode 0 = squeeze, 1 = Haar. */
int sininn_resample_nchw(const float* in, float* out, int B, int C, int H, int W,
                         int mode, int rev, float scale, sininn_stream_t stream);
/* Same maps on channels-last tensors: full-res [B,H,W,C] <-> [B,H/2,W/2,4C]. */
int sininn_resample_nhwc(const float* in, float* out, int B, int C, int H, int W,
                         int mode, int rev, float scale, sininn_stream_t stream);

/* Layout changes at the API boundary (the reference API is NCHW-contiguous,
 * loss.py:16-17).  chan_map (device int32[C], may be NULL): out channel i takes
 * in channel chan_map[i] -- folds FrEIA PermuteRandom (archs.py:65-68) into the
 * copy.  bf16_out (may be NULL): additionally writes out[:, c0:c1] as a compact
 * bf16 [npix][c1-c0] matrix (operand copy for the tensor-core subnets). */
int sininn_nchw_to_nhwc(const float* in, float* out, int B, int C, int HW, const int32_t* chan_map,
                        void* bf16_out, int c0, int c1, sininn_stream_t stream);
int sininn_nhwc_to_nchw(const float* in, float* out, int B, int C, int HW, const int32_t* chan_map,
                        sininn_stream_t stream);
/* Two IRevNetDownsampling nodes (mode 0 of sininn_resample_nchw, archs.py:28-38) followed by the NCHW -> NHWC change, as ONE
 * pass:  out[b][i][j][k2 * 4 C0 + k1 * C0 + c] = in[b][c][4 i + 2 (k2 >> 1) + (k1 >> 1)][4 j + 2 (k2 & 1) + (k1 & 1)]
 * (in [B][C0][H][W], out [B][H/4][W/4][16 C0]; H, W multiples of 4, C0 <= 8), bit-identical to the three separate calls.
 * bf16_out (may be NULL): compact bf16 copy of out channels [c0, c1).  sininn_nhwc_to_unsqueeze2 is the inverse map. */
int sininn_squeeze2_to_nhwc(const float* in, float* out, int B, int C0, int H, int W, void* bf16_out, int c0, int c1,
                            sininn_stream_t stream);
int sininn_nhwc_to_unsqueeze2(const float* in, float* out, int B, int C0, int H, int W, sininn_stream_t stream);

/* FrEIA PermuteRandom on channels-last data: out[p][i] = in[p][chan_map[i]] */
int sininn_permute_nhwc(const float* in, float* out, long long npix, int C, const int32_t* chan_map,
                        void* bf16_out, int c0, int c1, sininn_stream_t stream);

/* the same gather on two tensors of one shape in ONE launch (the backward pass undoes a permutation on the
 * activations and on their gradient with the same map); needs C % 4 == 0.  bf16_out_a (may be NULL): compact bf16
 * copy of out_a[:, c0:c1] (c0, c1 multiples of 4), the operand of the subnet evaluated next. */
int sininn_permute_nhwc_pair(const float* in_a, float* out_a, const float* in_b, float* out_b, long long npix, int C,
                             const int32_t* chan_map, void* bf16_out_a, int c0, int c1, sininn_stream_t stream);

/* Input pipeline (data.py:31-45 after the PNG decode): video [T][H][W][C] uint8 resident on the device, centers [B]
 * int32 frame indices; out [B][(2*win+1)*C][ph][pw] fp32 = frames centre-win..centre+win cropped to the patch at
 * (y0, x0), concatenated along channels, divided by 255.  win = 0, C = 3 gives the HR batch. */
int sininn_gather_windows_u8(const uint8_t* video, int T, int H, int W, int C, const int32_t* centers, int B, int win,
                             int y0, int x0, int ph, int pw, float* out, sininn_stream_t stream);

/* The same gather with one patch origin PER SAMPLE: crops_yx is a device int32 [B][2] of (y0, x0) on this video's
 * grid (the caller keeps y0 + ph <= H, x0 + pw <= W).  This is the "trained on patches" sampler the reference's
 * transform hook (data.py:43-44) leaves unused: every sample of a batch gets its own random crop. */
int sininn_gather_windows_u8_crops(const uint8_t* video, int T, int H, int W, int C, const int32_t* centers, int B, int win,
                                   const int32_t* crops_yx, int ph, int pw, float* out, sininn_stream_t stream);

/* Inference output path (lit_wrapper.py:117-121, transforms.ToPILImage on every frame): fp32 NCHW frames in [0, 1] ->
 * uint8 HWC, out[b][h][w][c] = (uint8) trunc(255 * clamp(in[b][c][h][w], 0, 1)).  For in-range values this is
 * pic.mul(255).byte(); out-of-range values are clamped (the reference's cast is undefined there). */
int sininn_quantize_u8_hwc(const float* in, uint8_t* out, int B, int C, int H, int W, sininn_stream_t stream);

/* Entry of the inverse pass: cat((lr, z), dim=1) of lit_wrapper.py:41-42 / 110-111 folded into the NCHW -> channels-last
 * change (plus the trailing PermuteRandom through chan_map and the bf16 operand copy, as sininn_nchw_to_nhwc).
 * lr [B][L][HW], z [B][Z][HW] fp32; out [B][HW][L+Z].  z == NULL: z = temp * N(0,1) is drawn on the device
 * (Philox4x32-10 keyed by `seed`, element i of the z tensor uses counter offset + i; Box-Muller), which replaces the
 * torch.randn + cat + copy of the reference; z_out (may be NULL) then receives the drawn z in NCHW.  step_ptr (may be
 * NULL): device int32 counter; the counter offset becomes offset + *step_ptr * step_stride, so a launch replayed from a
 * CUDA graph draws a new z whenever the counter (e.g. the optimizer's device-side step count) has advanced. */
int sininn_latent_to_nhwc(const float* lr, int L, const float* z, int Z, int B, int HW, const int32_t* chan_map, float* out,
                          void* bf16_out, int c0, int c1, unsigned long long seed, unsigned long long offset, float temp,
                          float* z_out, const int32_t* step_ptr, unsigned long long step_stride, sininn_stream_t stream);

/* FrEIA ActNorm (offered, commented out, at archs.py:40-44) on a channels-last fp32 matrix [npix][C], in place:
 *   inverse=0: u <- u * exp(log_scale[c]) + bias[c];   inverse=1: u <- (u - bias[c]) * exp(-log_scale[c]) */
int sininn_channel_affine(float* u, long long npix, int C, const float* log_scale, const float* bias, int inverse,
                          sininn_stream_t stream);
/* Its backward from the OUTPUT: on entry u = y, du = dL/dy of the direction `inverse` that produced y; on exit u = x,
 * du = dL/dx, and dlog_scale / dbias [C] (+)= the parameter gradients (deterministic two-pass sums; C <= 256). */
size_t sininn_channel_affine_bwd_workspace_bytes(void);
int sininn_channel_affine_bwd(float* u, float* du, long long npix, int C, const float* log_scale, const float* bias, int inverse,
                              float* dlog_scale, float* dbias, int accumulate, void* workspace, size_t workspace_bytes,
                              sininn_stream_t stream);
/* Log-determinant of a coupling half (FrEIA GLOWCouplingBlock.jacobian / last_jac): out[b] (+)= sign * sum over sample
 * b's pixels and L channels of g(s), g as in SININN_GLOW / SININN_IRN.  s [B * pix_per_sample][L], pixel stride s_stride. */
int sininn_logscale_sum(const float* s, int s_stride, int B, long long pix_per_sample, int L, int kind, float clamp, float sign,
                        float* out, int accumulate, sininn_stream_t stream);

/* ---------------------------------------------------------------- coupling
 * One half of an affine coupling, in place on a channel slice u[npix][L]
 * (pixel stride u_stride) of the fp32 trunk:
 *   inverse=0:  u <- exp(g(s)) * u + t        (GLOW y = e(s)*x + t; IRN y2 = x2*exp(s) + G, archs.py:154)
 *   inverse=1:  u <- (u - t) / exp(g(s))      (archs.py:157 and GLOW rev)
 * s,t: fp32 [npix][L] with their own strides (GLOW: both halves of one subnet
 * output; IRN: outputs of H and G).  u_bf16 (may be NULL): compact bf16 copy of
 * the updated slice.  fast_math = 1 (bf16 path): polynomial atan, ex2.approx and approximate division (errors at fp32
 * rounding level) instead of the accurate libm forms, whose ~60 instructions per element make the kernel ALU-bound. */
int sininn_coupling_apply(float* u, int u_stride, const float* s, int s_stride, const float* t, int t_stride,
                          long long npix, int L, int kind, float clamp, int inverse,
                          void* u_bf16, int fast_math, sininn_stream_t stream);
/* Backward of one coupling half from its OUTPUT (recompute-from-inverse, the
 * equations of SURVEY.md section 8a).  On entry u holds y and du holds dL/dy;
 * on exit u holds the reconstructed input x and du holds dL/dx.  Writes
 * dL/d(raw s) and dL/dt ([npix][L] each, dtype out_dtype, own strides).
 * `inverse` says which direction produced y (0: y=e*x+t, 1: y=(x-t)/e). */
int sininn_coupling_bwd(float* u, int u_stride, float* du, int du_stride,
                        const float* s, int s_stride, const float* t, int t_stride,
                        long long npix, int L, int kind, float clamp, int inverse,
                        void* ds_out, int ds_stride, void* dt_out, int dt_stride, int out_dtype,
                        void* x_bf16, int fast_math, sininn_stream_t stream);

/* The LAST half-step of a coupling block fused with the channel permutation that follows it (FrEIA GLOWCouplingBlock ->
 * PermuteRandom, archs.py:61-68):  out[p][i] = f(in[p][chan_map[i]]),  f = the half-step of sininn_coupling_apply for source
 * channels in [c0, c0 + L) (s, t indexed by channel - c0), identity for the others.  in / out: [npix][C] fp32 (C % 4 == 0),
 * not in place.  bf16_out (may be NULL): compact bf16 copy of out[:, bc0:bc1].  Same arithmetic as the two separate calls. */
int sininn_coupling_apply_permute(const float* in, float* out, long long npix, int C, const int32_t* chan_map, int c0, int L,
                                  const float* s, int s_stride, const float* t, int t_stride, int kind, float clamp, int inverse,
                                  void* bf16_out, int bc0, int bc1, int fast_math, sininn_stream_t stream);
/* The undo of that permutation fused with sininn_coupling_bwd of the half-step: y_in / dy_in are the trunk and its gradient in
 * the permuted layout, chan_map undoes the permutation (out channel i <- in channel chan_map[i]); output channels
 * [c0, c0 + L) (multiples of 4) receive x and dL/dx of the half-step, the others are moved unchanged; ds_out / dt_out as
 * in sininn_coupling_bwd. */
int sininn_coupling_bwd_unpermute(const float* y_in, const float* dy_in, float* x_out, float* dx_out, long long npix, int C,
                                  const int32_t* chan_map, int c0, int L, const float* s, int s_stride, const float* t, int t_stride,
                                  int kind, float clamp, int inverse, void* ds_out, int ds_stride, void* dt_out, int dt_stride,
                                  int out_dtype, int fast_math, sininn_stream_t stream);

/* small helpers around the subnets */
/* out[p][c] = scale * in[p][c] converted to out_dtype */
int sininn_cast_slice(const float* in, int in_stride, long long npix, int L, float scale, void* out, int out_dtype,
                      int out_stride, sininn_stream_t stream);
/* out <- d * act'(y) where y is the activation OUTPUT (sign-preserving activations); d fp32, y and out
 * in their own dtypes (out may alias d when out_dtype is fp32) */
int sininn_act_bwd(const float* d, int d_stride, const void* y, int y_dtype, int y_stride,
                   void* out, int out_dtype, int out_stride, long long npix, int L,
                   int act, float slope, sininn_stream_t stream);
/* out[j] (+)= sum_p in[p][j]; deterministic two-pass; workspace >= sininn_colsum_workspace_bytes */
size_t sininn_colsum_workspace_bytes(long long npix, int N);
int sininn_colsum(const void* in, int dtype, int in_stride, long long npix, int N, float* out, int accumulate,
                  void* workspace, size_t workspace_bytes, sininn_stream_t stream);
/* out (+)= alpha*a  over a [npix][L] slice (IRN additive half y1 = x1 +- F(x2), archs.py:152,158) */
int sininn_axpy_slice(float* out, int out_stride, const void* a, int a_dtype, int a_stride, long long npix, int L,
                      float alpha, sininn_stream_t stream);

/* ---------------------------------------------------------------- convolutions
 * Stride-1 "same" convolution with 1x1 or 3x3 taps on channels-last data as an
 * implicit GEMM  out[p][co] = sum_{tap,ci} in[p+off(tap)][ci] * wpack[tap][co][ci]
 * with a fused epilogue.  Replaces nn.Conv2d inside subnet_conv / subnet_conv_1x1
 * (archs.py:11-17) and DenseBlock (archs.py:77-81, 88-95); with dgrad-packed
 * weights the same call is the data gradient. */
typedef struct {
  int B, H, W;
  int Cin, Cout, taps;            /* taps = 1 or 9 */
  const void* in;  int in_dtype;  int in_stride;
  const void* wpack;              /* [taps][rows_pad][k_pad], dtype = in_dtype, zero padded */
  int rows_pad, k_pad;
  const float* bias;              /* [Cout] or NULL */
  void* out;       int out_dtype; int out_stride;
  int act; float slope;           /* activation applied to acc+bias */
  const void* mask; int mask_stride; /* NULL, or activation OUTPUT (dtype=out_dtype): result *= act'(mask) */
  int mask_act;
  int accumulate; float alpha;    /* out = (accumulate ? out : 0) + alpha * f(acc + bias) */
  /* tensor-core path only: 1 bit per element, uint32 [npix][ceil(Cout/32)]; channel c sits in word c/32 at bit
   * (e >> 1) + 16 * (e & 1), e = c % 32 (the two halves of a packed bf16x2 register map to bits j and 16 + j); an opaque
   * format between the call that writes bits_out and the calls that read it as mask_bits */
  const void* mask_bits;          /* NULL, or sign bits of the ReLU output the gradient flows through: result zeroed where 0 */
  void* bits_out;                 /* NULL, or receives the sign bits (value > 0) of this call's own output */
  /* Affine coupling fused into the epilogue (3x3 tensor-core path; cpl_mode 0 = off).  The convolution is the SECOND
   * conv of a GLOW subnet (archs.py:11-13 inside GLOWCouplingBlock, archs.py:61-64) packed with sininn_pack_conv_weight
   * mode 4 (output rows interleaved s_0, t_0, s_1, t_1, ...), Cout = 2 * cpl_L; its output [s | t] is consumed in
   * registers and never written (`out` may be NULL):
   *   cpl_mode 1: the half-step of sininn_coupling_apply on cpl_u [npix][cpl_L] (pixel stride cpl_u_stride), in place,
   *               direction cpl_inverse; cpl_bf16 (may be NULL) receives the compact bf16 copy of the result; cpl_a (may
   *               be NULL; fp32 [npix][2 * cpl_L], 16-byte aligned) receives the subnet output [s | t] (bias included,
   *               natural channel order) for a backward pass that keeps it instead of re-evaluating the subnet.
   *   cpl_mode 2: the half-step of sininn_coupling_bwd: cpl_u holds y -> x, cpl_du holds dL/dy -> dL/dx (in place),
   *               cpl_da (bf16 [npix][2 * cpl_L]) receives [dL/ds | dL/dt], cpl_bf16 the bf16 copy of x. */
  int cpl_mode; int cpl_L; int cpl_inverse; float cpl_clamp;
  float* cpl_u; int cpl_u_stride;
  float* cpl_du; int cpl_du_stride;
  void* cpl_bf16; void* cpl_da; float* cpl_a;
} sininn_conv_desc;

/* Debugging aid: when set to a device buffer of 4 x 512 int64, the CTA-pair 3x3 kernel records clock64() stamps of
 * CTA 0's producer / MMA / epilogue roles into words 0..1535 (only when built with -DSININN_PAIR_TRACE) and the CTA-pair
 * weight-gradient kernel its phase stamps into words 1536..1543: {entry, prologue done, dependency wait done, producer
 * done, first stage landed, last MMA issued, accumulators ready, partials stored}; the fused 1x1 subnet backward kernel
 * writes words 1600..1855 (MMA issuer, then epilogue warp 2: 16 tiles x 8 stamps each).  NULL switches tracing off (default). */
int sininn_debug_set_trace(void* device_buf_3x512_int64);

int sininn_conv_simt(const sininn_conv_desc* d, sininn_stream_t stream);   /* fp32-accurate CUDA-core path */
int sininn_conv_tc(const sininn_conv_desc* d, sininn_stream_t stream);     /* tcgen05/TMEM/TMA bf16 path */

/* Fused 1x1 coupling subnet, forward (tcgen05/TMEM/TMA, bf16 operands):
 *   out[p][:] = W2 * relu(W1 * x[p][:] + b1) + b2        per pixel p
 * Replaces subnet_conv_1x1 (archs.py:15-17: Conv2d(c_in,256,1) -> ReLU -> Conv2d(256,c_out,1)) as called by
 * the GLOW coupling halves (archs.py:56-64).  The hidden activation stays in shared memory; h_out / bits_out
 * (both optional) additionally store it (bf16 [npix][hidden]) and its ReLU sign bits (uint32 [npix][hidden/32])
 * for the backward kernels.  w1pack / w2pack are fprop packs (sininn_pack_conv_weight mode 0, taps = 1). */
typedef struct {
  long long npix;
  int Cin, hidden, Cout;
  const void* x;  int x_stride;          /* bf16 [npix][Cin] */
  const void* w1pack; int k1_pad;        /* bf16 [hidden][k1_pad] */
  const float* b1;                       /* [hidden] or NULL */
  const void* w2pack; int n2_pad;        /* bf16 [n2_pad][hidden] */
  const float* b2;                       /* [Cout] or NULL */
  float* out;     int out_stride;        /* fp32 [npix][Cout] */
  void* h_out;    int h_stride;          /* NULL or bf16 [npix][hidden] */
  void* bits_out;                        /* NULL or uint32 [npix][hidden/32] */
  /* The same pipeline is the subnet's DATA GRADIENT: with x = dL/d(out) [npix][Cout'], w1pack / w2pack the dgrad
   * packs (sininn_pack_conv_weight mode 1) of conv2 / conv1, b1 = b2 = NULL, mask_bits = the sign bits stored by
   * the forward call, the hidden tile is dL/dh = mask * (W2^T x) (h_out receives it for the weight gradient) and
   * out (+)= W1^T dL/dh is the gradient w.r.t. the subnet input. */
  const void* mask_bits;                 /* NULL, or uint32 [npix][hidden/32]: first stage = zero where the bit is 0 */
  int accumulate;                        /* 1: out += result */
  /* GLOW affine coupling in the second epilogue, exactly as the cpl_* fields of sininn_conv_desc: w2pack in the
   * interleaved mode-4 layout, Cout = 2 * cpl_L, `out` unused (may be NULL); cpl_mode 1 = half-step on cpl_u, cpl_mode 2 =
   * its backward (with h_out / bits_out stored for the subnet's backward pass, as in the recompute pass). */
  int cpl_mode; int cpl_L; int cpl_inverse; float cpl_clamp;
  float* cpl_u; int cpl_u_stride;
  float* cpl_du; int cpl_du_stride;
  void* cpl_bf16; void* cpl_da; float* cpl_a;
} sininn_subnet1x1_desc;

int sininn_subnet1x1_fwd_tc(const sininn_subnet1x1_desc* d, sininn_stream_t stream);
/* 1 when the fused kernel takes a Cin -> hidden -> Cout subnet (channel multiples and shared-memory budget), else 0 */
int sininn_subnet1x1_supported(int Cin, int hidden, int Cout);

/* Fused 1x1 coupling subnet, BACKWARD (tcgen05/TMEM/TMA, bf16 operands, hidden = 256): everything autograd derives for
 * subnet_conv_1x1 (archs.py:15-17) from x (the subnet input the forward pass read) and da = dL/d(subnet output), in one
 * kernel launch + one reduction launch:
 *   h = relu(W1 x + b1) (re-evaluated, bit-identical to sininn_subnet1x1_fwd_tc),  dh = (h > 0) * (W2^T da),
 *   dsrc += W1^T dh,   dw2 (+)= sum_p da h^T,  db2 (+)= sum_p da,   dw1 (+)= sum_p dh x^T,  db1 (+)= sum_p dh.
 * Neither h nor dh is written to memory.  w1pack: conv1 fprop pack (mode 0) [hidden][k1_pad]; w2dpack: conv2 dgrad pack
 * (mode 1) [hidden][k2_pad]; w1dpack: conv1 dgrad pack (mode 1) [r1_pad][hidden]; dw1 / dw2: OIHW fp32 ([hidden][Cin],
 * [Cout][hidden]).  Each CTA leaves one partial of the parameter gradients in `workspace`; they are summed in a fixed
 * order (bit-reproducible).  Shapes: sininn_subnet1x1_bwd_supported (Cin % 8 == 0, Cin <= 32, Cout % 4 == 0, Cout <= 64). */
typedef struct {
  long long npix;
  int Cin, hidden, Cout;
  const void* x;  int x_stride;          /* bf16 [npix][Cin] */
  const void* da; int da_stride;         /* bf16 [npix][Cout] */
  const void* w1pack;  int k1_pad;
  const float* b1;                       /* [hidden] or NULL */
  const void* w2dpack; int k2_pad;
  const void* w1dpack; int r1_pad;
  float* dsrc;    int dsrc_stride;       /* fp32 [npix][Cin], accumulated into */
  float* dw1; int dw1_accumulate; float* db1; int db1_accumulate;
  float* dw2; int dw2_accumulate; float* db2; int db2_accumulate;
  void* workspace; size_t workspace_bytes;
} sininn_subnet1x1_bwd_desc;

int sininn_subnet1x1_bwd_supported(int Cin, int hidden, int Cout);
size_t sininn_subnet1x1_bwd_workspace_bytes(const sininn_subnet1x1_bwd_desc* d);
int sininn_subnet1x1_bwd_tc(const sininn_subnet1x1_bwd_desc* d, sininn_stream_t stream);

/* Re-layout nn.Conv2d OIHW fp32 weights for the implicit GEMMs above.
 *   mode 0 (fprop): out[tap][co][ci] = w[co][ci][tap]           rows = Cout, k = Cin
 *   mode 1 (dgrad): out[tap][ci][co] = w[co][ci][taps-1-tap]    rows = Cin,  k = Cout
 *   mode 4 (fprop, rows interleaved): row 2c = output channel c, row 2c+1 = output channel Cout/2 + c -- the layout of a
 *               GLOW subnet's second convolution whose [s | t] halves are consumed pairwise by the fused coupling epilogue
 *   mode 2 / 3: the same two layouts for the fp32-accurate tensor-core path (bf16 only): K is six blocks of k_pad / 6
 *               holding [wh | wh | wm | wh | wm | wl] (wh = bf16(w), wm = bf16(w - wh), wl = bf16(w - wh - wm));
 *               pairs with sininn_split_bf16 operands */
int sininn_pack_conv_weight(const float* w_oihw, int Cout, int Cin, int taps, int mode,
                            void* out, int out_dtype, int rows_pad, int k_pad, sininn_stream_t stream);

/* The same re-layout for MANY weights in one launch.  jobs: device array of njobs x 8 int64
 * {src fp32 OIHW ptr, dst ptr, Cout, Cin, taps, mode, rows_pad, k_pad}; all outputs share out_dtype. */
int sininn_pack_conv_weights_batched(const void* jobs, int njobs, int out_dtype, sininn_stream_t stream);

/* fp32-accurate tensor-core path: an fp32 activation matrix [npix][L] (pixel stride in_stride) as a bf16 matrix
 * [npix][6 * Lp] of six channel blocks [h | m | h | l | m | h], h = bf16(s x), m = bf16(s x - h), l = bf16(s x - h - m)
 * (s = scale; Lp >= L a multiple of 8, padding columns zero).  sininn_conv_tc over this operand and a mode 2 / 3 pack
 * evaluates the fp32 convolution of archs.py:11-17 at fp32 accuracy on the bf16 tensor cores (6x the K: every product
 * term of the three-term expansions down to second order).  blocks = 4 writes only [h | m | h | l]: against the first
 * four K blocks of the same pack ([wh | wh | wm | wh]) that is the two-term-weight product, 2^-17 accurate at 4x the K --
 * enough wherever no ReLU decision hangs on the result (second convolution of a subnet, data gradients). */
int sininn_split_bf16(const float* in, int in_stride, long long npix, int L, float scale, void* out, int Lp, int blocks,
                      sininn_stream_t stream);
/* Weight gradient  dw[co][ci][tap] (+)= sum_p dy[p][co] * x[p+off(tap)][ci]  (OIHW fp32),
 * deterministic split over pixels + fixed-order reduction (no float atomics). */
typedef struct {
  int B, H, W;
  int Cin, Cout, taps;
  const void* x;  int x_dtype;  int x_stride;
  const void* dy; int dy_dtype; int dy_stride;
  float* dw;      int accumulate;
  void* workspace; size_t workspace_bytes;
  /* tensor-core path only: NULL, or the bias gradient  dbias[co] (+)= sum_p dy[p][co]  computed in the same two
   * launches (column sums of the dy tiles the weight-gradient kernel streams through shared memory anyway) */
  float* dbias;   int dbias_accumulate;
  /* fp32-accurate tensor-core path (CTA-pair kernel only; nterms = 0: off).  x and dy are split operands
   * (sininn_split_bf16): the gradient is accumulated over nterms pairs of channel blocks, pair t reading x channels
   * x_term_off[t] + [0, Cin) and dy channels dy_term_off[t] + [0, Cout) -- the products h*dh, h*dm, m*dh, m*dm, h*dl, l*dh
   * of the three-term expansions, i.e. the pixel (K) dimension is walked nterms times.  bias_term_mask: bit t set = pair
   * t's dy block enters the bias gradient (each of dh, dm, dl once). */
  int nterms; int x_term_off[6]; int dy_term_off[6]; int bias_term_mask;
  /* Several convolutions that read the SAME input as ONE problem (tensor-core path, CTA-pair kernel; nseg = 0: off) -- the
   * DenseBlock of archs.py:74-95, whose conv j reads the first cin_j channels of the growing concatenation x: dy holds their
   * output gradients side by side ([npix][Cout], Cout = sum of the segments' rows), Cin = the widest cin_j; dw / dbias are
   * unused.  Segment s owns dy channels [row0, row0 + rows) and receives  dw_s[rows][cin][taps] (+)= ...  (input channels
   * >= cin are computed and dropped) and, when dbias_s is not NULL,  dbias_s[rows] (+)= sum_p dy[p][row0 + r]. */
  int nseg;
  struct { int row0, rows, cin; float* dw; int accumulate; float* dbias; int dbias_accumulate; } seg[8];
} sininn_wgrad_desc;

/* 1 when the CTA-pair weight-gradient kernel takes this problem (the only kernel that serves nseg > 0) */
int sininn_wgrad_pair_supported(const sininn_wgrad_desc* d);

size_t sininn_wgrad_workspace_bytes(const sininn_wgrad_desc* d, int tensor_core);
int sininn_wgrad_simt(const sininn_wgrad_desc* d, sininn_stream_t stream);
int sininn_wgrad_tc(const sininn_wgrad_desc* d, sininn_stream_t stream);
/* Several weight gradients (e.g. the two convolutions of a coupling subnet, archs.py:11-17, whose operands become
 * available together in the backward pass) in ONE pair of launches: the grid is divided among the problems in
 * proportion to their work, so each is cut in fewer pixel splits than alone.  `workspace` serves the whole group
 * (>= sininn_wgrad_group_workspace_bytes); the workspace fields of the descriptors are ignored.  Results are
 * bit-identical from call to call for the same group, but not to the one-by-one calls (different split counts). */
size_t sininn_wgrad_group_workspace_bytes(const sininn_wgrad_desc* descs, int n);
int sininn_wgrad_tc_group(const sininn_wgrad_desc* descs, int n, void* workspace, size_t workspace_bytes,
                          sininn_stream_t stream);

/* ---------------------------------------------------------------- caller-side fusions
 * sum((a[:, :L]-b)^2) over a strided slice -> out[0] (+ optional gradient 2*scale*(a-b)); used for
 * loss.reconstruction (loss.py:3-5) and latent_nll (loss.py:38-39, b = NULL). */
size_t sininn_sqdiff_workspace_bytes(long long n);
int sininn_sqdiff_nchw(const float* a, const float* b, long long n, float scale, float* loss_out,
                       float* grad_out, void* workspace, size_t workspace_bytes, sininn_stream_t stream);
/* The forward half's loss in one pass (lit_wrapper.py:45-48 with loss.py:3-5,38-39): y [B][C][HW] is the network output,
 * lr [B][L][HW] the low-resolution target;  loss = w_rec * mean((y[:, :L] - lr)^2) + w_nll * mean(y[:, L:]^2)  and
 * grad_out [B][C][HW] = dloss/dy.  workspace >= 2 * sininn_sqdiff_workspace_bytes(B*C*HW). */
int sininn_inn_fwd_loss(const float* y, const float* lr, int B, int C, int L, long long HW, float w_rec, float w_nll,
                        float* loss_out, float* grad_out, void* workspace, size_t workspace_bytes, sininn_stream_t stream);
/* loss.mmd (loss.py:9-36): multi-kernel inverse-multiquadric MMD between the flattened batches x, y [b][D] (b <= 64),
 * kernels (0.2,2),(1.5,2),(3.0,2) or, rev != 0, (0.2,0.1),(0.2,0.5),(0.2,2).  loss_out[0] = scale * mmd; grad_x_out
 * (may be NULL) [b][D] = scale * d mmd / d x.  Device-agnostic restatement (the reference hard-codes .to('cuda')):
 * Gram matrices by a split over D with a fixed-order second pass, so the value is bit-reproducible. */
size_t sininn_mmd_workspace_bytes(int b, long long D);
int sininn_mmd(const float* x, const float* y, int b, long long D, int rev, float scale, float* loss_out, float* grad_x_out,
               void* workspace, size_t workspace_bytes, sininn_stream_t stream);

/* torch.optim.Adam semantics (lit_wrapper.py:134-137: L2 weight decay folded into the gradient) over a
 * flat fp32 arena; step is the 1-based step count. */
int sininn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                     float grad_scale, sininn_stream_t stream);
/* The same update with the step count kept on the DEVICE so that the launch can be replayed from a CUDA graph:
 * step_state is 3 x 4 bytes {int32 steps done so far, float scratch, float scratch}, zero-initialised by the
 * caller; every call increments it and derives the bias corrections from it (two launches). */
int sininn_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int* step_state,
                         float grad_scale, sininn_stream_t stream);

/* The same with the gradient given as the sum of TWO arenas (grad + grad_b, grad_b may be NULL): the two halves of a
 * training step accumulate into separate arenas on separate streams; adding them inside Adam saves a pass. */
int sininn_adam_step_dev2(float* param, const float* grad, const float* grad_b, float* exp_avg, float* exp_avg_sq, long long n,
                          float lr, float beta1, float beta2, float eps, float weight_decay, int* step_state,
                          float grad_scale, sininn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SININN_H */
