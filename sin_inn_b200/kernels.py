"""Thin typed wrappers: torch tensors / 2-D strided views -> libsininn C-ABI calls.

Every function launches on torch's current CUDA stream and returns immediately.
A "view" is a 2-D fp32/bf16 tensor [npix, L] with unit channel stride and an
arbitrary row stride (a channel slice of a wider channels-last matrix).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import F32, BF16, ConvDesc, Subnet1x1BwdDesc, Subnet1x1Desc, WgradDesc, check, dtype_code, load, stream_ptr

_workspaces = {}
SPLIT_BLOCKS = 6          # channel blocks of a split operand / K blocks of a split weight pack (conv_simt.cu)

# ---- launch accounting / optional CUDA-event profiling (used by bench.py; off by default)
LAUNCHES = 0          # kernels launched through this module since last reset
_prof = None


def profile_begin():
    """Start timing every launch with CUDA events on the launching stream, grouped by kernel family."""
    global _prof
    _prof = []


def profile_end(by_shape=False):
    """Stop profiling; returns {family: {"ms", "flops", "bytes", "n"}} (algorithmic flops/bytes as passed in).
    by_shape: key on "family shape-tag" (the tag the launch wrappers attach) instead of the family alone."""
    global _prof
    rec, _prof = _prof or [], None
    torch.cuda.synchronize()
    out = {}
    for fam, e0, e1, nk, flops, nbytes, tag in rec:
        d = out.setdefault(f"{fam} {tag}" if (by_shape and tag) else fam, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += flops
        d["bytes"] += nbytes
        d["n"] += nk
    return out


def _run(family, fn, nk=1, flops=0.0, nbytes=0.0, tag=None):
    global LAUNCHES
    LAUNCHES += nk
    if _prof is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    _prof.append((family, e0, e1, nk, flops, nbytes, tag() if callable(tag) else tag))
    return rc


def _workspace(dev, name, nbytes):
    # one buffer per (device, purpose, STREAM): the two halves of a training step run on two streams
    key = (dev.index, name, torch.cuda.current_stream(dev).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


def _view2d(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.SininnError(f"expected a 2-D view with unit channel stride, got shape {tuple(t.shape)} stride {t.stride()}")
    _lib.require_cuda(t)
    return t


def _p(t):
    return 0 if t is None else t.data_ptr()


# ----------------------------------------------------------------------------- resampling / layout
def resample_nchw(x, mode, rev, scale=1.0):
    """mode 0 squeeze / 1 Haar on NCHW fp32.  rev=0: [B,C,H,W]->[B,4C,H/2,W/2]."""
    _lib.require_cuda(x)
    x = x.contiguous()
    B, c, h, w = x.shape
    if not rev:
        C_, H, W = c, h, w
        out = torch.empty(B, 4 * c, h // 2, w // 2, dtype=torch.float32, device=x.device)
    else:
        if c % 4:
            raise _lib.SininnError(f"inverse resample needs channels % 4 == 0, got {c}")
        C_, H, W = c // 4, 2 * h, 2 * w
        out = torch.empty(B, c // 4, 2 * h, 2 * w, dtype=torch.float32, device=x.device)
    check(_run("resample", lambda: load().sininn_resample_nchw(x.data_ptr(), out.data_ptr(), B, C_, H, W, mode, int(rev), float(scale),
                                      stream_ptr()), 1, 0.0, 8.0 * x.numel()), "resample_nchw")
    return out


def resample_nhwc(x, mode, rev, scale=1.0):
    """Same maps on channels-last [B,H,W,C] fp32 tensors."""
    _lib.require_cuda(x)
    B, h, w, c = x.shape
    if not rev:
        C_, H, W = c, h, w
        out = torch.empty(B, h // 2, w // 2, 4 * c, dtype=torch.float32, device=x.device)
    else:
        C_, H, W = c // 4, 2 * h, 2 * w
        out = torch.empty(B, 2 * h, 2 * w, c // 4, dtype=torch.float32, device=x.device)
    check(_run("resample", lambda: load().sininn_resample_nhwc(x.data_ptr(), out.data_ptr(), B, C_, H, W, mode, int(rev), float(scale),
                                      stream_ptr()), 1, 0.0, 8.0 * x.numel()), "resample_nhwc")
    return out


def nchw_to_nhwc(x, chan_map=None, bf16_range=None):
    _lib.require_cuda(x)
    x = x.contiguous()
    B, c, h, w = x.shape
    out = torch.empty(B, h, w, c, dtype=torch.float32, device=x.device)
    bf = None
    c0 = c1 = 0
    if bf16_range is not None:
        c0, c1 = bf16_range
        bf = torch.empty(B * h * w, c1 - c0, dtype=torch.bfloat16, device=x.device)
    check(_run("layout", lambda: load().sininn_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, c, h * w, _p(chan_map), _p(bf), c0, c1,
                                     stream_ptr()), 1, 0.0, 8.0 * x.numel()), "nchw_to_nhwc")
    return out, bf


def latent_to_nhwc(lr, z, z_dims, chan_map=None, bf16_range=None, seed=0, offset=0, temp=1.0, z_out=None, step_state=None):
    """cat((lr, z), 1) -> channels-last [B, h, w, L+Z] in one pass (lit_wrapper.py:41-42).  z=None: z = temp*N(0,1)
    is drawn inside the kernel (counter offset + step_state[0] * z.numel() when a device step counter is given)."""
    _lib.require_cuda(lr)
    lr = lr.contiguous()
    B, L, h, w = lr.shape
    if z is not None:
        z = z.contiguous()
        if tuple(z.shape) != (B, z_dims, h, w):
            raise _lib.SininnError(f"latent_to_nhwc: z must be {(B, z_dims, h, w)}, got {tuple(z.shape)}")
    Cc = L + z_dims
    out = torch.empty(B, h, w, Cc, dtype=torch.float32, device=lr.device)
    bf, c0, c1 = None, 0, 0
    if bf16_range is not None:
        c0, c1 = bf16_range
        bf = torch.empty(B * h * w, c1 - c0, dtype=torch.bfloat16, device=lr.device)
    stride = B * z_dims * h * w
    check(_run("layout", lambda: load().sininn_latent_to_nhwc(lr.data_ptr(), L, _p(z), z_dims, B, h * w, _p(chan_map), out.data_ptr(), _p(bf),
                                                              c0, c1, int(seed) & (2**64 - 1), int(offset), float(temp), _p(z_out),
                                                              _p(step_state), stride, stream_ptr()), 1, 0.0, 8.0 * out.numel()),
          "latent_to_nhwc")
    return out, bf


def channel_affine(U, log_scale, bias, inverse):
    """ActNorm on a channels-last fp32 tensor [..., C], in place."""
    c = U.shape[-1]
    check(_run("misc", lambda: load().sininn_channel_affine(U.data_ptr(), U.numel() // c, c, log_scale.data_ptr(), bias.data_ptr(),
                                                            int(inverse), stream_ptr()), 1, 0.0, 8.0 * U.numel()), "channel_affine")
    return U


def channel_affine_bwd(U, dU, log_scale, bias, inverse, dls, dbias, accumulate):
    """U: y -> x, dU: dy -> dx in place; dls / dbias [C] (+)= parameter gradients."""
    c = U.shape[-1]
    lib = load()
    ws = _workspace(U.device, "affine_bwd", lib.sininn_channel_affine_bwd_workspace_bytes())
    check(_run("misc", lambda: lib.sininn_channel_affine_bwd(U.data_ptr(), dU.data_ptr(), U.numel() // c, c, log_scale.data_ptr(),
                                                             bias.data_ptr(), int(inverse), dls.data_ptr(), dbias.data_ptr(), int(accumulate),
                                                             ws.data_ptr(), ws.numel(), stream_ptr()), 2, 0.0, 16.0 * U.numel()),
          "channel_affine_bwd")


def logscale_sum(s, B, kind, clamp, sign, out, accumulate):
    """out[b] (+)= sign * sum_{pixels of sample b, channels} g(s);  s: [npix, L] fp32 view."""
    s = _view2d(s)
    npix, L = s.shape
    check(_run("misc", lambda: load().sininn_logscale_sum(s.data_ptr(), s.stride(0), B, npix // B, L, kind, float(clamp), float(sign),
                                                          out.data_ptr(), int(accumulate), stream_ptr()), 1), "logscale_sum")
    return out


def mmd(x, y, rev, scale, want_grad=False):
    """scale * loss.mmd(x, y, rev) (loss.py:9-36) -> (0-dim loss, gradient w.r.t. x or None); x, y [b, ...] fp32."""
    _lib.require_cuda(x)
    x, y = x.contiguous(), y.contiguous()
    b = x.shape[0]
    D = x.numel() // b
    if y.numel() != x.numel():
        raise _lib.SininnError("mmd: x and y must have the same number of elements")
    lib = load()
    nbytes = lib.sininn_mmd_workspace_bytes(b, D)
    if nbytes == 0:
        raise _lib.SininnError(f"mmd: unsupported batch size {b} (1..64)")
    ws = _workspace(x.device, "mmd", nbytes)
    out = torch.empty((), dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    check(_run("loss", lambda: lib.sininn_mmd(x.data_ptr(), y.data_ptr(), b, D, int(bool(rev)), float(scale), out.data_ptr(), _p(grad),
                                              ws.data_ptr(), ws.numel(), stream_ptr()), 4 if want_grad else 3), "mmd")
    return out, grad


def nhwc_to_nchw(x, chan_map=None):
    _lib.require_cuda(x)
    B, h, w, c = x.shape
    out = torch.empty(B, c, h, w, dtype=torch.float32, device=x.device)
    check(_run("layout", lambda: load().sininn_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), B, c, h * w, _p(chan_map), stream_ptr()), 1, 0.0, 8.0 * x.numel()),
          "nhwc_to_nchw")
    return out


def squeeze2_to_nhwc(x, bf16_range=None):
    """Two squeezes (resample_nchw mode 0) + nchw_to_nhwc in one pass: [B,C0,H,W] -> ([B,H/4,W/4,16*C0], bf16 copy or None)."""
    _lib.require_cuda(x)
    B, c0, H, W = x.shape
    out = torch.empty(B, H // 4, W // 4, 16 * c0, dtype=torch.float32, device=x.device)
    bf, b0, b1 = None, 0, 0
    if bf16_range is not None:
        b0, b1 = bf16_range
        bf = torch.empty(B * (H // 4) * (W // 4), b1 - b0, dtype=torch.bfloat16, device=x.device)
    check(_run("layout", lambda: load().sininn_squeeze2_to_nhwc(x.data_ptr(), out.data_ptr(), B, c0, H, W, _p(bf), b0, b1, stream_ptr()), 1, 0.0,
               8.0 * x.numel()), "squeeze2_to_nhwc")
    return out, bf


def nhwc_to_unsqueeze2(x):
    """Inverse of squeeze2_to_nhwc: [B,h,w,16*C0] -> [B,C0,4h,4w]."""
    _lib.require_cuda(x)
    B, h, w, c = x.shape
    out = torch.empty(B, c // 16, 4 * h, 4 * w, dtype=torch.float32, device=x.device)
    check(_run("layout", lambda: load().sininn_nhwc_to_unsqueeze2(x.data_ptr(), out.data_ptr(), B, c // 16, 4 * h, 4 * w, stream_ptr()), 1, 0.0,
               8.0 * x.numel()), "nhwc_to_unsqueeze2")
    return out


def permute_nhwc(x, chan_map, bf16_range=None):
    """x: channels-last tensor [..., C]; out[..., i] = x[..., chan_map[i]]."""
    _lib.require_cuda(x)
    c = x.shape[-1]
    npix = x.numel() // c
    out = torch.empty_like(x)
    bf = None
    c0 = c1 = 0
    if bf16_range is not None:
        c0, c1 = bf16_range
        bf = torch.empty(npix, c1 - c0, dtype=torch.bfloat16, device=x.device)
    check(_run("permute", lambda: load().sininn_permute_nhwc(x.data_ptr(), out.data_ptr(), npix, c, chan_map.data_ptr(), _p(bf), c0, c1,
                                     stream_ptr()), 1, 0.0, 8.0 * x.numel()), "permute_nhwc")
    return out, bf


def gather_windows_u8(video, centers, win, crop=None, crops_yx=None, patch=None):
    """video: uint8 [T, H, W, C] on the device; centers: int32 [B] on the device; crop = (y0, x0, ph, pw) or None.
    Returns fp32 [B, (2*win+1)*C, ph, pw] = the frame windows / 255 (data.py:31-45 without the PNG decode).
    crops_yx (int32 [B, 2] on the device) + patch = (ph, pw): one patch origin per sample instead of one crop."""
    _lib.require_cuda(video)
    if video.dtype != torch.uint8 or video.dim() != 4 or not video.is_contiguous():
        raise _lib.SininnError("gather_windows_u8: video must be a contiguous uint8 [T, H, W, C] tensor")
    T, H, W, Cc = video.shape
    B = centers.numel()
    if crops_yx is not None:
        ph, pw = patch
        if crops_yx.dtype != torch.int32 or tuple(crops_yx.shape) != (B, 2) or not crops_yx.is_contiguous():
            raise _lib.SininnError("gather_windows_u8: crops_yx must be a contiguous int32 [B, 2] tensor")
        out = torch.empty(B, (2 * win + 1) * Cc, ph, pw, dtype=torch.float32, device=video.device)
        check(_run("misc", lambda: load().sininn_gather_windows_u8_crops(video.data_ptr(), T, H, W, Cc, centers.data_ptr(), B, int(win),
                                                                         crops_yx.data_ptr(), int(ph), int(pw), out.data_ptr(), stream_ptr()),
                   1, 0.0, 5.0 * out.numel()), "gather_windows_u8_crops")
        return out
    y0, x0, ph, pw = crop if crop is not None else (0, 0, H, W)
    out = torch.empty(B, (2 * win + 1) * Cc, ph, pw, dtype=torch.float32, device=video.device)
    check(_run("misc", lambda: load().sininn_gather_windows_u8(video.data_ptr(), T, H, W, Cc, centers.data_ptr(), B, int(win), int(y0),
                                          int(x0), int(ph), int(pw), out.data_ptr(), stream_ptr()), 1, 0.0, 5.0 * out.numel()),
          "gather_windows_u8")
    return out


def quantize_u8_hwc(x):
    """fp32 NCHW frames in [0, 1] -> uint8 [B, H, W, C] (ToPILImage's mul(255).byte(), clamped), on the device."""
    _lib.require_cuda(x)
    x = x.contiguous()
    B, c, h, w = x.shape
    out = torch.empty(B, h, w, c, dtype=torch.uint8, device=x.device)
    check(_run("misc", lambda: load().sininn_quantize_u8_hwc(x.data_ptr(), out.data_ptr(), B, c, h, w, stream_ptr()), 1, 0.0,
               5.0 * x.numel()), "quantize_u8_hwc")
    return out


def permute_nhwc_pair(xa, xb, chan_map, bf16_range=None):
    """Both gathers of the backward pass (activations and their gradient, same map) in one launch; optionally also
    the compact bf16 copy of the gathered activations' channels bf16_range.  Returns (out_a, out_b, bf16 or None)."""
    _lib.require_cuda(xa)
    c = xa.shape[-1]
    if c % 4 or xa.shape != xb.shape or (bf16_range is not None and (bf16_range[0] % 4 or bf16_range[1] % 4)):
        oa, bf = permute_nhwc(xa, chan_map, bf16_range)
        return oa, permute_nhwc(xb, chan_map)[0], bf
    npix = xa.numel() // c
    oa, ob = torch.empty_like(xa), torch.empty_like(xb)
    bf, c0, c1 = None, 0, 0
    if bf16_range is not None:
        c0, c1 = bf16_range
        bf = torch.empty(npix, c1 - c0, dtype=torch.bfloat16, device=xa.device)
    check(_run("permute", lambda: load().sininn_permute_nhwc_pair(xa.data_ptr(), oa.data_ptr(), xb.data_ptr(), ob.data_ptr(), npix, c,
                                          chan_map.data_ptr(), _p(bf), c0, c1, stream_ptr()), 1, 0.0, 16.0 * xa.numel()), "permute_nhwc_pair")
    return oa, ob, bf


# ----------------------------------------------------------------------------- coupling
def coupling_apply(u, s, t, kind, clamp, inverse, want_bf16=False, fast=False):
    u, s, t = _view2d(u), _view2d(s), _view2d(t)
    npix, L = u.shape
    bf = torch.empty(npix, L, dtype=torch.bfloat16, device=u.device) if want_bf16 else None
    check(_run("coupling", lambda: load().sininn_coupling_apply(u.data_ptr(), u.stride(0), s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0),
                                       npix, L, kind, float(clamp), int(inverse), _p(bf), int(fast), stream_ptr()), 1, 0.0, 16.0 * npix * L),
          "coupling_apply")
    return bf


def coupling_bwd(u, du, s, t, kind, clamp, inverse, ds, dt, want_bf16=False, fast=False):
    u, du, s, t, ds, dt = map(_view2d, (u, du, s, t, ds, dt))
    npix, L = u.shape
    if ds.dtype != dt.dtype:
        raise _lib.SininnError("coupling_bwd: ds and dt must share a dtype")
    bf = torch.empty(npix, L, dtype=torch.bfloat16, device=u.device) if want_bf16 else None
    check(_run("coupling_bwd", lambda: load().sininn_coupling_bwd(u.data_ptr(), u.stride(0), du.data_ptr(), du.stride(0), s.data_ptr(), s.stride(0),
                                     t.data_ptr(), t.stride(0), npix, L, kind, float(clamp), int(inverse),
                                     ds.data_ptr(), ds.stride(0), dt.data_ptr(), dt.stride(0), dtype_code(ds), _p(bf), int(fast),
                                     stream_ptr()), 1, 0.0, (24.0 + 2.0 * ds.element_size()) * npix * L), "coupling_bwd")
    return bf


def coupling_apply_permute(U, chan_map, rng, s, t, kind, clamp, inverse, bf16_range=None, fast=False):
    """Last half-step of a coupling block + the channel permutation after it in one pass: returns (U_new, bf16 copy of
    U_new[..., bf16_range] or None) with U_new[..., i] = f(U[..., chan_map[i]]), f = the half-step on source channels rng."""
    _lib.require_cuda(U)
    C_ = U.shape[-1]
    npix = U.numel() // C_
    s, t = _view2d(s), _view2d(t)
    out = torch.empty_like(U)
    bf, b0, b1 = None, 0, 0
    if bf16_range is not None:
        b0, b1 = bf16_range
        bf = torch.empty(npix, b1 - b0, dtype=torch.bfloat16, device=U.device)
    c0, L = rng[0], rng[1] - rng[0]
    check(_run("coupling", lambda: load().sininn_coupling_apply_permute(U.data_ptr(), out.data_ptr(), npix, C_, chan_map.data_ptr(), c0, L,
                                       s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0), kind, float(clamp), int(inverse), _p(bf), b0, b1,
                                       int(fast), stream_ptr()), 1, 0.0, 8.0 * U.numel() + 8.0 * npix * L), "coupling_apply_permute")
    return out, bf


def coupling_bwd_unpermute(Y, dY, chan_map, rng, s, t, kind, clamp, inverse, ds, dt, fast=False):
    """Undo of a channel permutation on (trunk, gradient) + the backward of the half-step on channels rng of the result, one
    pass: returns (X, dX) in the un-permuted layout; ds / dt receive the half-step's subnet-output gradients."""
    _lib.require_cuda(Y)
    C_ = Y.shape[-1]
    npix = Y.numel() // C_
    s, t, ds, dt = map(_view2d, (s, t, ds, dt))
    if ds.dtype != dt.dtype:
        raise _lib.SininnError("coupling_bwd_unpermute: ds and dt must share a dtype")
    X, dX = torch.empty_like(Y), torch.empty_like(dY)
    c0, L = rng[0], rng[1] - rng[0]
    check(_run("coupling_bwd", lambda: load().sininn_coupling_bwd_unpermute(Y.data_ptr(), dY.data_ptr(), X.data_ptr(), dX.data_ptr(), npix, C_,
                                       chan_map.data_ptr(), c0, L, s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0), kind, float(clamp),
                                       int(inverse), ds.data_ptr(), ds.stride(0), dt.data_ptr(), dt.stride(0), dtype_code(ds), int(fast),
                                       stream_ptr()), 1, 0.0, 16.0 * Y.numel() + (8.0 + 2.0 * ds.element_size()) * npix * L), "coupling_bwd_unpermute")
    return X, dX


def cast_slice(src, out, scale=1.0):
    """out[p][c] = scale*src[p][c]; src fp32 view, out fp32/bf16 view of the same shape."""
    src, out = _view2d(src), _view2d(out)
    npix, L = src.shape
    check(_run("misc", lambda: load().sininn_cast_slice(src.data_ptr(), src.stride(0), npix, L, float(scale), out.data_ptr(),
                                   dtype_code(out), out.stride(0), stream_ptr()), 1, 0.0, (4.0 + out.element_size()) * npix * L), "cast_slice")
    return out


def split_bf16(src, scale=1.0, blocks=None):
    """fp32 view [npix, L] -> bf16 [npix, 6*Lp] of the channel blocks [h | m | h | l | m | h] (Lp = L rounded up to 8):
    the operand form of the fp32-accurate tensor-core path (sininn_split_bf16)."""
    src = _view2d(src)
    npix, L = src.shape
    Lp = (L + 7) // 8 * 8
    blocks = SPLIT_BLOCKS if blocks is None else blocks       # 4: [h | m | h | l] only (two-term-weight product)
    out = torch.empty(npix, blocks * Lp, dtype=torch.bfloat16, device=src.device)
    check(_run("split", lambda: load().sininn_split_bf16(src.data_ptr(), src.stride(0), npix, L, float(scale), out.data_ptr(), Lp,
                                                         blocks, stream_ptr()), 1, 0.0, (4.0 + 2.0 * blocks) * npix * L), "split_bf16")
    return out


def act_bwd(d, y, out, act, slope=0.0):
    """out = d * act'(y); d fp32 view, y the activation output, out fp32/bf16 view (may be d itself)."""
    d, y, out = _view2d(d), _view2d(y), _view2d(out)
    npix, L = d.shape
    check(_run("misc", lambda: load().sininn_act_bwd(d.data_ptr(), d.stride(0), y.data_ptr(), dtype_code(y), y.stride(0), out.data_ptr(),
                                dtype_code(out), out.stride(0), npix, L, act, float(slope), stream_ptr()), 1, 0.0, 0.0), "act_bwd")
    return out


def colsum(src, out, accumulate=False):
    src = _view2d(src)
    npix, N = src.shape
    lib = load()
    nbytes = lib.sininn_colsum_workspace_bytes(npix, N)
    ws = _workspace(src.device, "colsum", nbytes)
    check(_run("colsum", lambda: lib.sininn_colsum(src.data_ptr(), dtype_code(src), src.stride(0), npix, N,
                                                   out.data_ptr(), int(accumulate), ws.data_ptr(), ws.numel(),
                                                   stream_ptr()), 2, 0.0, float(src.element_size()) * npix * N), "colsum")
    return out


def axpy_slice(out, a, alpha):
    out, a = _view2d(out), _view2d(a)
    npix, L = out.shape
    check(_run("misc", lambda: load().sininn_axpy_slice(out.data_ptr(), out.stride(0), a.data_ptr(), dtype_code(a), a.stride(0), npix, L,
                                   float(alpha), stream_ptr()), 1, 0.0, 0.0), "axpy_slice")


# ----------------------------------------------------------------------------- convolutions
def pack_weight(w, mode, dtype, rows_pad, k_pad):
    """w: OIHW fp32 parameter.  mode 0 fprop [tap][co][ci]; mode 1 dgrad [tap][ci][co] (flipped)."""
    _lib.require_cuda(w, "weight")
    co, ci, kh, kw = w.shape
    taps = kh * kw
    out = torch.empty(taps, rows_pad, k_pad, dtype=dtype, device=w.device)
    wc = w.detach().contiguous()
    check(_run("pack", lambda: load().sininn_pack_conv_weight(wc.data_ptr(), co, ci, taps, mode, out.data_ptr(), dtype_code(out), rows_pad,
                                         k_pad, stream_ptr()), 1, 0.0, 0.0), "pack_conv_weight")
    return out


def pack_weights_batched(jobs, njobs, dtype):
    """jobs: device int64 tensor [njobs, 8] = {src, dst, Cout, Cin, taps, mode, rows_pad, k_pad}."""
    code = F32 if dtype == torch.float32 else BF16
    check(_run("pack", lambda: load().sininn_pack_conv_weights_batched(jobs.data_ptr(), njobs, code, stream_ptr()), 1),
          "pack_conv_weights_batched")


def conv(x, wpack, geom, cout, out, bias=None, act=0, slope=0.0, mask=None, mask_act=0, accumulate=False, alpha=1.0,
         tensor_core=False, mask_bits=None, bits_out=None, coupling=None):
    """Implicit-GEMM 1x1 / 3x3 convolution on channels-last views.
    x: [npix, Cin] view; wpack: [taps, rows_pad, k_pad]; geom = (B, H, W); out: [npix, cout] view.
    coupling (3x3 tensor-core path, wpack in the interleaved mode-4 layout, out = None): the GLOW affine coupling runs in the
    epilogue on the subnet output held in registers -- dict(mode=1|2, u=, clamp=, inverse=, bf16=None, du=None, da=None),
    see sininn_conv_desc."""
    x = _view2d(x)
    if coupling is not None:
        return _conv_coupled(x, wpack, geom, cout, bias, coupling)
    out = _view2d(out)
    B, H, W = geom
    d = ConvDesc()
    d.B, d.H, d.W = B, H, W
    d.Cin, d.Cout, d.taps = x.shape[1], cout, wpack.shape[0]
    d.inp, d.in_dtype, d.in_stride = x.data_ptr(), dtype_code(x), x.stride(0)
    if wpack.dtype != x.dtype:
        raise _lib.SininnError("conv: packed weights and input must share a dtype")
    d.wpack, d.rows_pad, d.k_pad = wpack.data_ptr(), wpack.shape[1], wpack.shape[2]
    d.bias = _p(bias)
    d.out, d.out_dtype, d.out_stride = out.data_ptr(), dtype_code(out), out.stride(0)
    d.act, d.slope = act, float(slope)
    if mask is not None:
        mask = _view2d(mask)
        if mask.dtype != out.dtype:
            raise _lib.SininnError("conv: mask dtype must equal out dtype")
        d.mask, d.mask_stride, d.mask_act = mask.data_ptr(), mask.stride(0), mask_act
    else:
        d.mask, d.mask_stride, d.mask_act = 0, 0, 0
    d.accumulate, d.alpha = int(accumulate), float(alpha)
    if (mask_bits is not None or bits_out is not None) and not tensor_core:
        raise _lib.SininnError("conv: sign-bit masks are a tensor-core-path feature")
    d.mask_bits, d.bits_out = _p(mask_bits), _p(bits_out)
    lib = load()
    fn = lib.sininn_conv_tc if tensor_core else lib.sininn_conv_simt
    flops = 2.0 * B * H * W * d.Cin * d.Cout * d.taps
    fam = "conv3x3" if d.taps == 9 else "conv1x1"
    tag = lambda: (f"{H}x{W} {d.Cin}->{d.Cout}" + (" relu" if act else "") + (" bits" if bits_out is not None else "")
                   + (" masked" if (mask is not None or mask_bits is not None) else "") + (" acc" if accumulate else "")
                   + (" f32" if out.dtype == torch.float32 else " bf16"))
    check(_run(fam, lambda: fn(C.byref(d), stream_ptr()), 1, flops, tag=tag), "conv_tc" if tensor_core else "conv_simt")
    return out


def _conv_coupled(x, wpack, geom, cout, bias, cp):
    B, H, W = geom
    u = _view2d(cp["u"])
    L = u.shape[1]
    d = ConvDesc()
    d.B, d.H, d.W = B, H, W
    d.Cin, d.Cout, d.taps = x.shape[1], cout, wpack.shape[0]
    d.inp, d.in_dtype, d.in_stride = x.data_ptr(), dtype_code(x), x.stride(0)
    d.wpack, d.rows_pad, d.k_pad = wpack.data_ptr(), wpack.shape[1], wpack.shape[2]
    d.bias = _p(bias)
    d.out, d.out_dtype, d.out_stride = 0, F32, 0
    d.act, d.slope, d.mask, d.mask_stride, d.mask_act = 0, 0.0, 0, 0, 0
    d.accumulate, d.alpha, d.mask_bits, d.bits_out = 0, 1.0, 0, 0
    d.cpl_mode, d.cpl_L, d.cpl_inverse, d.cpl_clamp = int(cp["mode"]), L, int(bool(cp["inverse"])), float(cp["clamp"])
    d.cpl_u, d.cpl_u_stride = u.data_ptr(), u.stride(0)
    du = cp.get("du")
    if du is not None:
        du = _view2d(du)
        d.cpl_du, d.cpl_du_stride = du.data_ptr(), du.stride(0)
    else:
        d.cpl_du, d.cpl_du_stride = 0, 0
    d.cpl_bf16, d.cpl_da, d.cpl_a = _p(cp.get("bf16")), _p(cp.get("da")), _p(cp.get("a"))
    flops = 2.0 * B * H * W * d.Cin * d.Cout * d.taps
    tag = lambda: f"{H}x{W} {d.Cin}->{d.Cout} cpl{d.cpl_mode}" + (" +a" if d.cpl_a else "")
    check(_run("conv3x3", lambda: load().sininn_conv_tc(C.byref(d), stream_ptr()), 1, flops, tag=tag), "conv_tc(coupling)")


def subnet1x1_supported(cin, hidden, cout):
    """Shapes the fused 1x1 subnet kernel takes (everything else runs as two conv launches)."""
    return bool(load().sininn_subnet1x1_supported(int(cin), int(hidden), int(cout)))


def subnet1x1_fwd(x, w1pack, b1, w2pack, b2, out, h_out=None, bits_out=None, mask_bits=None, accumulate=False, coupling=None):
    """out = W2 relu(W1 x + b1) + b2 per pixel, one launch (tcgen05; hidden activation stays on chip).
    x: bf16 [npix, Cin] view; w1pack [1, hidden, k1_pad]; w2pack [1, n2_pad, hidden]; out: fp32 [npix, Cout] view;
    h_out (bf16 [npix, hidden]) / bits_out (int32 [npix, hidden/32]) optionally receive the hidden activation."""
    x = _view2d(x)
    d = Subnet1x1Desc()
    d.npix = x.shape[0]
    if coupling is not None:             # GLOW half-step in the second epilogue: w2pack interleaved (mode 4), out unused
        u = _view2d(coupling["u"])
        cout = 2 * u.shape[1]
        d.cpl_mode, d.cpl_L, d.cpl_inverse, d.cpl_clamp = int(coupling["mode"]), u.shape[1], int(bool(coupling["inverse"])), float(coupling["clamp"])
        d.cpl_u, d.cpl_u_stride = u.data_ptr(), u.stride(0)
        du = coupling.get("du")
        if du is not None:
            du = _view2d(du)
            d.cpl_du, d.cpl_du_stride = du.data_ptr(), du.stride(0)
        else:
            d.cpl_du, d.cpl_du_stride = 0, 0
        d.cpl_bf16, d.cpl_da, d.cpl_a = _p(coupling.get("bf16")), _p(coupling.get("da")), _p(coupling.get("a"))
    else:
        out = _view2d(out)
        cout = out.shape[1]
        d.cpl_mode = 0
    d.Cin, d.hidden, d.Cout = x.shape[1], w1pack.shape[1], cout
    d.x, d.x_stride = x.data_ptr(), x.stride(0)
    d.w1pack, d.k1_pad = w1pack.data_ptr(), w1pack.shape[2]
    d.b1 = _p(b1)
    d.w2pack, d.n2_pad = w2pack.data_ptr(), w2pack.shape[1]
    d.b2 = _p(b2)
    d.out, d.out_stride = (out.data_ptr(), out.stride(0)) if coupling is None else (0, 0)
    if x.dtype != torch.bfloat16 or (coupling is None and out.dtype != torch.float32) or w1pack.dtype != torch.bfloat16 or w2pack.dtype != torch.bfloat16:
        raise _lib.SininnError("subnet1x1_fwd: bf16 operands and an fp32 output are required")
    if w2pack.shape[2] != d.hidden or w1pack.shape[0] != 1 or w2pack.shape[0] != 1:
        raise _lib.SininnError("subnet1x1_fwd: packed weights do not describe a 1x1 Cin->hidden->Cout subnet")
    if h_out is not None:
        h_out = _view2d(h_out)
        d.h_out, d.h_stride = h_out.data_ptr(), h_out.stride(0)
    else:
        d.h_out, d.h_stride = 0, 0
    d.bits_out = _p(bits_out)
    # gradient use (see sininn.h): hidden stage masked by the forward's ReLU sign bits, output accumulated
    d.mask_bits, d.accumulate = _p(mask_bits), int(accumulate)
    flops = 2.0 * d.npix * d.hidden * (d.Cin + d.Cout)
    tag = lambda: (f"npix{d.npix} {d.Cin}->{d.hidden}->{d.Cout}" + (" keep" if d.h_out else "") + (" grad" if d.mask_bits else "")
                   + (f" cpl{d.cpl_mode}" if d.cpl_mode else ""))
    check(_run("subnet1x1", lambda: load().sininn_subnet1x1_fwd_tc(C.byref(d), stream_ptr()), 1, flops, tag=tag), "subnet1x1_fwd_tc")
    return out


def subnet1x1_bwd_supported(cin, hidden, cout):
    """Shapes the fused 1x1 subnet BACKWARD kernel takes (sininn_subnet1x1_bwd_tc)."""
    return bool(load().sininn_subnet1x1_bwd_supported(int(cin), int(hidden), int(cout)))


def subnet1x1_bwd(x, da, w1pack, b1, w2dpack, w1dpack, dsrc, grads1, grads2):
    """Whole backward pass of a 1x1 subnet a = W2 relu(W1 x + b1) + b2 in one kernel (+ one reduction) launch: the hidden
    activation is re-evaluated from x on chip, dsrc += W1^T dh, and the parameter gradients of both convolutions.
    x: bf16 [npix, Cin] view; da: bf16 [npix, Cout] view; w1pack / w2dpack / w1dpack: conv1 fprop, conv2 dgrad, conv1 dgrad
    packs; dsrc: fp32 [npix, Cin] view; grads1 / grads2 = (dw, accumulate, dbias, dbias_accumulate) of conv1 / conv2."""
    x, da, dsrc = _view2d(x), _view2d(da), _view2d(dsrc)
    d = Subnet1x1BwdDesc()
    d.npix = x.shape[0]
    d.Cin, d.hidden, d.Cout = x.shape[1], w1pack.shape[1], da.shape[1]
    if (x.dtype != torch.bfloat16 or da.dtype != torch.bfloat16 or dsrc.dtype != torch.float32
            or any(w.dtype != torch.bfloat16 or w.shape[0] != 1 for w in (w1pack, w2dpack, w1dpack))):
        raise _lib.SininnError("subnet1x1_bwd: bf16 operands / packs (1 tap) and an fp32 input gradient are required")
    if w2dpack.shape[1] != d.hidden or w1dpack.shape[2] != d.hidden or dsrc.shape[1] != d.Cin or da.shape[0] != d.npix:
        raise _lib.SininnError("subnet1x1_bwd: packed weights / operands do not describe a 1x1 Cin->hidden->Cout subnet")
    d.x, d.x_stride = x.data_ptr(), x.stride(0)
    d.da, d.da_stride = da.data_ptr(), da.stride(0)
    d.w1pack, d.k1_pad = w1pack.data_ptr(), w1pack.shape[2]
    d.b1 = _p(b1)
    d.w2dpack, d.k2_pad = w2dpack.data_ptr(), w2dpack.shape[2]
    d.w1dpack, d.r1_pad = w1dpack.data_ptr(), w1dpack.shape[1]
    d.dsrc, d.dsrc_stride = dsrc.data_ptr(), dsrc.stride(0)
    (dw1, acc1, db1, accb1), (dw2, acc2, db2, accb2) = grads1, grads2
    for t in (dw1, db1, dw2, db2):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.SininnError("subnet1x1_bwd: gradient tensors must be contiguous fp32")
    d.dw1, d.dw1_accumulate, d.db1, d.db1_accumulate = dw1.data_ptr(), int(acc1), db1.data_ptr(), int(accb1)
    d.dw2, d.dw2_accumulate, d.db2, d.db2_accumulate = dw2.data_ptr(), int(acc2), db2.data_ptr(), int(accb2)
    lib = load()
    ws = _workspace(x.device, "s1bwd", lib.sininn_subnet1x1_bwd_workspace_bytes(C.byref(d)))
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    # recompute of h (Cin) + dh (Cout) + dsrc (Cin) + both weight gradients (Cin + Cout)
    flops = 2.0 * d.npix * d.hidden * (3 * d.Cin + 2 * d.Cout)
    tag = lambda: f"npix{d.npix} {d.Cin}->{d.hidden}->{d.Cout}"
    check(_run("subnet1x1_bwd", lambda: lib.sininn_subnet1x1_bwd_tc(C.byref(d), stream_ptr()), 2, flops, tag=tag), "subnet1x1_bwd_tc")


def wgrad(x, dy, geom, taps, dw, accumulate=False, tensor_core=False, dbias=None, dbias_accumulate=False):
    """dw[co][ci][tap] (+)= sum_p dy[p][co] * x[p+off(tap)][ci];  dw: OIHW fp32 contiguous.
    dbias (tensor-core path only): additionally dbias[co] (+)= sum_p dy[p][co] in the same launches."""
    x, dy = _view2d(x), _view2d(dy)
    B, H, W = geom
    d = WgradDesc()
    d.B, d.H, d.W = B, H, W
    d.Cin, d.Cout, d.taps = x.shape[1], dy.shape[1], taps
    d.x, d.x_dtype, d.x_stride = x.data_ptr(), dtype_code(x), x.stride(0)
    d.dy, d.dy_dtype, d.dy_stride = dy.data_ptr(), dtype_code(dy), dy.stride(0)
    d.dw, d.accumulate = dw.data_ptr(), int(accumulate)
    if dbias is not None and not tensor_core:
        raise _lib.SininnError("wgrad: the fused bias gradient is a tensor-core-path feature (use colsum)")
    d.dbias, d.dbias_accumulate = _p(dbias), int(dbias_accumulate)
    d.nterms = 0
    lib = load()
    nbytes = lib.sininn_wgrad_workspace_bytes(C.byref(d), int(tensor_core))
    ws = _workspace(x.device, "wgrad", nbytes)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
    fn = lib.sininn_wgrad_tc if tensor_core else lib.sininn_wgrad_simt
    flops = 2.0 * B * H * W * d.Cin * d.Cout * d.taps
    check(_run("wgrad", lambda: fn(C.byref(d), stream_ptr()), 2, flops), "wgrad_tc" if tensor_core else "wgrad_simt")
    return dw


def wgrad_group(jobs):
    """Several tensor-core weight (+ bias) gradients in ONE pair of launches (sininn_wgrad_tc_group).
    jobs: list of (x, dy, geom, taps, dw, accumulate, dbias, dbias_accumulate[, terms]) as for wgrad(); terms (split
    operands of the fp32-accurate path) = (x block offsets, dy block offsets, bias mask), see sininn_wgrad_desc."""
    n = len(jobs)
    arr = (WgradDesc * n)()
    flops = 0.0
    for d, job in zip(arr, jobs):
        x, dy, geom, taps, dw, acc, dbias, dbacc = job[:8]
        terms = job[8] if len(job) > 8 else None
        d.nterms = 0
        if terms is not None:
            xo, yo, mask = terms
            d.nterms, d.bias_term_mask = len(xo), int(mask)
            for t in range(len(xo)):
                d.x_term_off[t], d.dy_term_off[t] = int(xo[t]), int(yo[t])
        x, dy = _view2d(x), _view2d(dy)
        d.B, d.H, d.W = geom
        d.Cin, d.Cout, d.taps = x.shape[1], dy.shape[1], taps
        d.x, d.x_dtype, d.x_stride = x.data_ptr(), dtype_code(x), x.stride(0)
        d.dy, d.dy_dtype, d.dy_stride = dy.data_ptr(), dtype_code(dy), dy.stride(0)
        d.dw, d.accumulate = dw.data_ptr(), int(acc)
        d.dbias, d.dbias_accumulate = _p(dbias), int(dbacc)
        d.workspace, d.workspace_bytes = 0, 0
        flops += 2.0 * geom[0] * geom[1] * geom[2] * d.Cin * d.Cout * taps * max(1, d.nterms)
    lib = load()
    ws = _workspace(jobs[0][0].device, "wgrad", lib.sininn_wgrad_group_workspace_bytes(arr, n))
    tag = lambda: " | ".join(f"{a.H}x{a.W} {a.Cin}x{a.Cout} t{a.taps}" for a in arr)
    check(_run("wgrad", lambda: lib.sininn_wgrad_tc_group(arr, n, ws.data_ptr(), ws.numel(), stream_ptr()), 2, flops, tag=tag),
          "wgrad_tc_group")


def _merged_desc(x, dy, geom, taps, segments):
    d = WgradDesc()
    d.B, d.H, d.W = geom
    d.Cin, d.Cout, d.taps = x.shape[1], dy.shape[1], taps
    d.x, d.x_dtype, d.x_stride = x.data_ptr(), dtype_code(x), x.stride(0)
    d.dy, d.dy_dtype, d.dy_stride = dy.data_ptr(), dtype_code(dy), dy.stride(0)
    d.dw, d.accumulate, d.dbias, d.dbias_accumulate, d.nterms = 0, 0, 0, 0, 0
    d.workspace, d.workspace_bytes = 0, 0
    d.nseg = len(segments)
    for sg, (row0, rows, cin, dw, acc, dbias, dbacc) in zip(d.seg, segments):
        sg.row0, sg.rows, sg.cin = int(row0), int(rows), int(cin)
        sg.dw, sg.accumulate = dw.data_ptr(), int(acc)
        sg.dbias, sg.dbias_accumulate = _p(dbias), int(dbacc)
    return d


def wgrad_merged_supported(x, dy, geom, taps):
    """Does the CTA-pair weight-gradient kernel take the merged problem (all convolutions of a DenseBlock at once)?"""
    if x.dtype != torch.bfloat16 or dy.dtype != torch.bfloat16 or x.stride(0) % 8 or dy.stride(0) % 8:
        return False
    d = _merged_desc(x, dy, geom, taps, [])
    return bool(load().sininn_wgrad_pair_supported(C.byref(d)))


def wgrad_merged(x, dy, geom, taps, segments):
    """Weight (+ bias) gradients of several convolutions that read the same input as ONE tensor-core problem (the DenseBlock
    of archs.py:74-95): x [npix, Cin] the concatenation, dy [npix, sum rows] the output gradients side by side; segments =
    [(row0, rows, cin, dw, accumulate, dbias or None, dbias_accumulate)]: dw_s[rows][cin][taps] (+)= dy[:, row0:row0+rows]^T x[:, :cin]."""
    x, dy = _view2d(x), _view2d(dy)
    if len(segments) > 8:
        raise _lib.SininnError("wgrad_merged: at most 8 segments")
    arr = (WgradDesc * 1)()
    arr[0] = _merged_desc(x, dy, geom, taps, segments)
    lib = load()
    ws = _workspace(x.device, "wgrad", lib.sininn_wgrad_group_workspace_bytes(arr, 1))
    flops = sum(2.0 * geom[0] * geom[1] * geom[2] * rows * cin * taps for _, rows, cin, *_ in segments)
    tag = lambda: f"{geom[1]}x{geom[2]} merged {x.shape[1]}x{dy.shape[1]} t{taps} ({len(segments)} convs)"
    check(_run("wgrad", lambda: lib.sininn_wgrad_tc_group(arr, 1, ws.data_ptr(), ws.numel(), stream_ptr()), 2, flops, tag=tag),
          "wgrad_tc_group(merged)")


# ----------------------------------------------------------------------------- caller-side fusions
def sqdiff(a, b, scale, want_grad=False):
    """scale * sum((a-b)^2) -> 0-dim tensor; optional gradient 2*scale*(a-b)."""
    _lib.require_cuda(a)
    a = a.contiguous()
    n = a.numel()
    lib = load()
    ws = _workspace(a.device, "sqdiff", lib.sininn_sqdiff_workspace_bytes(n))
    out = torch.empty((), dtype=torch.float32, device=a.device)
    grad = torch.empty_like(a) if want_grad else None
    if b is not None:
        b = b.contiguous()
    check(_run("loss", lambda: lib.sininn_sqdiff_nchw(a.data_ptr(), _p(b), n, float(scale), out.data_ptr(), _p(grad),
                                                      ws.data_ptr(), ws.numel(), stream_ptr()), 2), "sqdiff")
    return out, grad


def inn_fwd_loss(y, lr, w_rec, w_nll):
    """w_rec*mean((y[:, :L]-lr)^2) + w_nll*mean(y[:, L:]^2) -> (0-dim loss, d loss / d y), one pass over y."""
    _lib.require_cuda(y)
    y, lr = y.contiguous(), lr.contiguous()
    B, Cc, H, W = y.shape
    L = lr.shape[1]
    lib = load()
    ws = _workspace(y.device, "fwd_loss", 2 * lib.sininn_sqdiff_workspace_bytes(y.numel()))
    out = torch.empty((), dtype=torch.float32, device=y.device)
    grad = torch.empty_like(y)
    check(_run("loss", lambda: lib.sininn_inn_fwd_loss(y.data_ptr(), lr.data_ptr(), B, Cc, L, H * W, float(w_rec), float(w_nll),
                                                       out.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()), 2),
          "inn_fwd_loss")
    return out, grad


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, step, grad_scale=1.0):
    n = param.numel()
    check(_run("adam", lambda: load().sininn_adam_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), n,
                                  float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                  int(step), float(grad_scale), stream_ptr()), 1, 0.0, 28.0 * n), "adam_step")


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, step_state, grad_scale=1.0, grad_b=None):
    """Adam with the step count on the device (int32[3] state tensor): replayable from a CUDA graph.
    grad_b: optional second gradient arena, added to grad inside the kernel."""
    n = param.numel()
    check(_run("adam", lambda: load().sininn_adam_step_dev2(param.data_ptr(), grad.data_ptr(), _p(grad_b), exp_avg.data_ptr(),
                                      exp_avg_sq.data_ptr(), n, float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                      step_state.data_ptr(), float(grad_scale), stream_ptr()), 2, 0.0, 28.0 * n + (4.0 * n if grad_b is not None else 0.0)),
          "adam_step_dev")
