"""ctypes binding of libsininn.so (the C ABI declared in include/sininn.h).

There is NO fallback: if the shared library is missing or a call fails, an
exception is raised.  PyTorch is used only to own device memory and streams.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsininn.so")

F32, BF16 = 0, 1
GLOW, IRN = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2

_c_ll = C.c_longlong
_vp = C.c_void_p


class SininnError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("Cin", C.c_int), ("Cout", C.c_int), ("taps", C.c_int),
        ("inp", _vp), ("in_dtype", C.c_int), ("in_stride", C.c_int),
        ("wpack", _vp), ("rows_pad", C.c_int), ("k_pad", C.c_int),
        ("bias", _vp),
        ("out", _vp), ("out_dtype", C.c_int), ("out_stride", C.c_int),
        ("act", C.c_int), ("slope", C.c_float),
        ("mask", _vp), ("mask_stride", C.c_int), ("mask_act", C.c_int),
        ("accumulate", C.c_int), ("alpha", C.c_float),
        ("mask_bits", _vp), ("bits_out", _vp),
        ("cpl_mode", C.c_int), ("cpl_L", C.c_int), ("cpl_inverse", C.c_int), ("cpl_clamp", C.c_float),
        ("cpl_u", _vp), ("cpl_u_stride", C.c_int),
        ("cpl_du", _vp), ("cpl_du_stride", C.c_int),
        ("cpl_bf16", _vp), ("cpl_da", _vp), ("cpl_a", _vp),
    ]


class WgradSegment(C.Structure):
    _fields_ = [("row0", C.c_int), ("rows", C.c_int), ("cin", C.c_int), ("dw", _vp), ("accumulate", C.c_int),
                ("dbias", _vp), ("dbias_accumulate", C.c_int)]


class WgradDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("Cin", C.c_int), ("Cout", C.c_int), ("taps", C.c_int),
        ("x", _vp), ("x_dtype", C.c_int), ("x_stride", C.c_int),
        ("dy", _vp), ("dy_dtype", C.c_int), ("dy_stride", C.c_int),
        ("dw", _vp), ("accumulate", C.c_int),
        ("workspace", _vp), ("workspace_bytes", C.c_size_t),
        ("dbias", _vp), ("dbias_accumulate", C.c_int),
        ("nterms", C.c_int), ("x_term_off", C.c_int * 6), ("dy_term_off", C.c_int * 6), ("bias_term_mask", C.c_int),
        ("nseg", C.c_int), ("seg", WgradSegment * 8),
    ]


class Subnet1x1Desc(C.Structure):
    _fields_ = [
        ("npix", _c_ll),
        ("Cin", C.c_int), ("hidden", C.c_int), ("Cout", C.c_int),
        ("x", _vp), ("x_stride", C.c_int),
        ("w1pack", _vp), ("k1_pad", C.c_int),
        ("b1", _vp),
        ("w2pack", _vp), ("n2_pad", C.c_int),
        ("b2", _vp),
        ("out", _vp), ("out_stride", C.c_int),
        ("h_out", _vp), ("h_stride", C.c_int),
        ("bits_out", _vp),
        ("mask_bits", _vp), ("accumulate", C.c_int),
        ("cpl_mode", C.c_int), ("cpl_L", C.c_int), ("cpl_inverse", C.c_int), ("cpl_clamp", C.c_float),
        ("cpl_u", _vp), ("cpl_u_stride", C.c_int),
        ("cpl_du", _vp), ("cpl_du_stride", C.c_int),
        ("cpl_bf16", _vp), ("cpl_da", _vp), ("cpl_a", _vp),
    ]


class Subnet1x1BwdDesc(C.Structure):
    _fields_ = [
        ("npix", _c_ll),
        ("Cin", C.c_int), ("hidden", C.c_int), ("Cout", C.c_int),
        ("x", _vp), ("x_stride", C.c_int),
        ("da", _vp), ("da_stride", C.c_int),
        ("w1pack", _vp), ("k1_pad", C.c_int),
        ("b1", _vp),
        ("w2dpack", _vp), ("k2_pad", C.c_int),
        ("w1dpack", _vp), ("r1_pad", C.c_int),
        ("dsrc", _vp), ("dsrc_stride", C.c_int),
        ("dw1", _vp), ("dw1_accumulate", C.c_int), ("db1", _vp), ("db1_accumulate", C.c_int),
        ("dw2", _vp), ("dw2_accumulate", C.c_int), ("db2", _vp), ("db2_accumulate", C.c_int),
        ("workspace", _vp), ("workspace_bytes", C.c_size_t),
    ]


# name -> (restype, argtypes); every symbol include/sininn.h declares
SIGNATURES = {
    "sininn_version": (C.c_int, []),
    "sininn_last_error": (C.c_char_p, []),
    "sininn_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "sininn_resample_nchw": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _vp]),
    "sininn_resample_nhwc": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _vp]),
    "sininn_nchw_to_nhwc": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "sininn_nhwc_to_nchw": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "sininn_squeeze2_to_nhwc": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "sininn_nhwc_to_unsqueeze2": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "sininn_permute_nhwc": (C.c_int, [_vp, _vp, _c_ll, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "sininn_permute_nhwc_pair": (C.c_int, [_vp, _vp, _vp, _vp, _c_ll, C.c_int, _vp, _vp, C.c_int, C.c_int, _vp]),
    "sininn_gather_windows_u8": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, _vp, _vp]),
    "sininn_gather_windows_u8_crops": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int,
                                                 _vp, _vp]),
    "sininn_latent_to_nhwc": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int,
                                        C.c_ulonglong, C.c_ulonglong, C.c_float, _vp, _vp, C.c_ulonglong, _vp]),
    "sininn_channel_affine": (C.c_int, [_vp, _c_ll, C.c_int, _vp, _vp, C.c_int, _vp]),
    "sininn_channel_affine_bwd_workspace_bytes": (C.c_size_t, []),
    "sininn_channel_affine_bwd": (C.c_int, [_vp, _vp, _c_ll, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_size_t, _vp]),
    "sininn_logscale_sum": (C.c_int, [_vp, C.c_int, C.c_int, _c_ll, C.c_int, C.c_int, C.c_float, C.c_float, _vp, C.c_int, _vp]),
    "sininn_mmd_workspace_bytes": (C.c_size_t, [C.c_int, _c_ll]),
    "sininn_mmd": (C.c_int, [_vp, _vp, C.c_int, _c_ll, C.c_int, C.c_float, _vp, _vp, _vp, C.c_size_t, _vp]),
    "sininn_quantize_u8_hwc": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "sininn_coupling_apply": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _c_ll, C.c_int, C.c_int, C.c_float,
                                        C.c_int, _vp, C.c_int, _vp]),
    "sininn_coupling_bwd": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _c_ll, C.c_int, C.c_int,
                                      C.c_float, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]),
    "sininn_coupling_apply_permute": (C.c_int, [_vp, _vp, _c_ll, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int,
                                                C.c_float, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "sininn_coupling_bwd_unpermute": (C.c_int, [_vp, _vp, _vp, _vp, _c_ll, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int,
                                                C.c_int, C.c_float, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "sininn_cast_slice": (C.c_int, [_vp, C.c_int, _c_ll, C.c_int, C.c_float, _vp, C.c_int, C.c_int, _vp]),
    "sininn_act_bwd": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _c_ll, C.c_int, C.c_int,
                                 C.c_float, _vp]),
    "sininn_colsum_workspace_bytes": (C.c_size_t, [_c_ll, C.c_int]),
    "sininn_colsum": (C.c_int, [_vp, C.c_int, C.c_int, _c_ll, C.c_int, _vp, C.c_int, _vp, C.c_size_t, _vp]),
    "sininn_axpy_slice": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, _c_ll, C.c_int, C.c_float, _vp]),
    "sininn_debug_set_trace": (C.c_int, [_vp]),
    "sininn_conv_simt": (C.c_int, [C.POINTER(ConvDesc), _vp]),
    "sininn_conv_tc": (C.c_int, [C.POINTER(ConvDesc), _vp]),
    "sininn_subnet1x1_fwd_tc": (C.c_int, [C.POINTER(Subnet1x1Desc), _vp]),
    "sininn_subnet1x1_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "sininn_subnet1x1_bwd_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "sininn_subnet1x1_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(Subnet1x1BwdDesc)]),
    "sininn_subnet1x1_bwd_tc": (C.c_int, [C.POINTER(Subnet1x1BwdDesc), _vp]),
    "sininn_pack_conv_weight": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "sininn_pack_conv_weights_batched": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "sininn_split_bf16": (C.c_int, [_vp, C.c_int, _c_ll, C.c_int, C.c_float, _vp, C.c_int, C.c_int, _vp]),
    "sininn_wgrad_workspace_bytes": (C.c_size_t, [C.POINTER(WgradDesc), C.c_int]),
    "sininn_wgrad_simt": (C.c_int, [C.POINTER(WgradDesc), _vp]),
    "sininn_wgrad_tc": (C.c_int, [C.POINTER(WgradDesc), _vp]),
    "sininn_wgrad_pair_supported": (C.c_int, [C.POINTER(WgradDesc)]),
    "sininn_wgrad_group_workspace_bytes": (C.c_size_t, [C.POINTER(WgradDesc), C.c_int]),
    "sininn_wgrad_tc_group": (C.c_int, [C.POINTER(WgradDesc), C.c_int, _vp, C.c_size_t, _vp]),
    "sininn_sqdiff_workspace_bytes": (C.c_size_t, [_c_ll]),
    "sininn_sqdiff_nchw": (C.c_int, [_vp, _vp, _c_ll, C.c_float, _vp, _vp, _vp, C.c_size_t, _vp]),
    "sininn_inn_fwd_loss": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _c_ll, C.c_float, C.c_float, _vp, _vp, _vp, C.c_size_t, _vp]),
    "sininn_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _c_ll, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                   C.c_int, C.c_float, _vp]),
    "sininn_adam_step_dev": (C.c_int, [_vp, _vp, _vp, _vp, _c_ll, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                       _vp, C.c_float, _vp]),
    "sininn_adam_step_dev2": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _c_ll, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                        _vp, C.c_float, _vp]),
}

_lib = None


def load():
    """Load libsininn.so; raises (never falls back) if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SininnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C sin_inn_b200/csrc`). There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SininnError(f"libsininn.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sininn_last_error().decode("utf-8", "replace")
        raise SininnError(f"{what or 'libsininn'} failed (code {rc}): {msg}")


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise SininnError(f"unsupported dtype {t.dtype}")


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise SininnError(f"{what} must be a CUDA tensor: this package has no CPU path "
                          f"(the CPU oracle lives in oracle/ and is test infrastructure only)")
