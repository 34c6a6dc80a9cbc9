"""Plan executor for invertible networks on the libsininn kernels.

A network is compiled into a list of plan ops (resample / coupling block /
channel permutation).  The executor keeps the activations ("trunk") as ONE fp32
channels-last tensor that every coupling half-step updates in place, and runs

  * forward            net(x)            (reference: FrEIA ReversibleGraphNet.forward,
  * inverse            net(x, rev=True)   call site archs.py:71; InvRescaleNet.forward archs.py:223-233)
  * backward           from the OUTPUT only: each block's input is reconstructed with
                       the exact inverse while gradients flow,
                       so no activation is stored between forward and backward.

The only tensors that cross the API are NCHW-contiguous fp32 (loss.py:16-17 needs
`.view(b, c*h*w)` to work on the result).
"""
import os
import weakref
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import kernels as K
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, GLOW, IRN, SininnError, require_cuda


@dataclass(frozen=True)
class EngineConfig:
    # "bf16":   bf16 operands, fp32 accumulate (tcgen05 kernels; the headline path)
    # "fp32":   CUDA-core fp32 subnets: the reference-accurate path (1e-4 parity on outputs AND gradients)
    # "fp32tc": fp32 activations, subnet GEMMs on the tensor cores over bf16 hi/mid/lo split operands: outputs 1e-5,
    #           round trip 2e-6, ~6.5x the speed of "fp32"; gradients are limited by ReLU-kink flips
    precision: str = "bf16"
    tensor_core: bool = True     # tcgen05 kernels for the subnets (bf16 / fp32tc)
    # What a differentiable pass keeps for its backward.  The TRUNK is never kept: it is rebuilt block by block from the
    # exact inverse.  The coupling subnets' internals (operand copy, hidden activation, ReLU sign bits, output) cannot be
    # rebuilt from the inverse, only re-evaluated:
    #   "recompute": re-evaluate every subnet during backward (memory of one block at a time),
    #   "store":     keep them from the value pass (what autograd does in the reference; ~100 B per pixel and subnet at the
    #                headline shape, 3.3 GB per step at B=32 256x256 -- against 180 GB), no forward GEMMs in backward,
    #   "auto":      store when the estimate fits in STORE_FRACTION of the memory the process has not allocated.
    activations: str = "auto"

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.precision == "bf16" else torch.float32

    @property
    def tc(self):
        return self.tensor_core and self.precision == "bf16"

    @property
    def split(self):
        """fp32-accurate path on the tensor cores: fp32 activations / gradients, every subnet GEMM evaluated on bf16
        hi/mid/lo splits of its operands (kernels.split_bf16, pack modes 2 / 3)."""
        return self.tensor_core and self.precision == "fp32tc"


PRECISIONS = ("bf16", "fp32", "fp32tc")
ACTIVATION_MODES = ("auto", "store", "recompute")
STORE_FRACTION = float(os.environ.get("SININN_STORE_FRACTION", "0.25"))


def default_config():
    prec = os.environ.get("SININN_PRECISION", "bf16")
    if prec not in PRECISIONS:
        raise SininnError(f"SININN_PRECISION must be one of {PRECISIONS}, got {prec!r}")
    tc = os.environ.get("SININN_TENSOR_CORE", "1") != "0"
    keep = os.environ.get("SININN_ACTIVATIONS", "auto")
    if keep not in ACTIVATION_MODES:
        raise SininnError(f"SININN_ACTIVATIONS must be one of {ACTIVATION_MODES}, got {keep!r}")
    return EngineConfig(precision=prec, tensor_core=tc, activations=keep)


def _round_up(v, m):
    return (v + m - 1) // m * m


FUSE_1X1 = os.environ.get("SININN_FUSE_1X1", "1") != "0"     # fused 1x1 subnet kernel (subnet1x1_tc.cu)
# fused 1x1 subnet BACKWARD kernel (subnet1x1_bwd.cu): hidden activation re-evaluated on chip, data gradient and both
# weight / bias gradients in one launch -- the value pass then keeps only the subnet's bf16 operand copy
FUSE_1X1_BWD = os.environ.get("SININN_FUSE_1X1_BWD", "1") != "0"
# GLOW affine coupling (value pass / backward pass) inside the epilogue of the 3x3 subnet's second convolution
# bit 0: value passes; bit 1: backward passes; bit 2: backward passes only where the subnet output is <= 64 columns wide
# (level 0: that convolution is bound by its small-N MMAs and its epilogue has slack; at level 1 the extra loads, stores
# and transcendental math of the backward half-step make the epilogue warps the bottleneck)
FUSE_COUPLING = int(os.environ.get("SININN_FUSE_COUPLING", "5"))
FUSE_COUPLING_1X1 = int(os.environ.get("SININN_FUSE_COUPLING_1X1", "1"))     # the same bits for the fused 1x1 subnet kernel
# "store" mode value passes: 1 = fused epilogue that also writes the subnet output for backward (cpl_a).  Measured on the
# B200 at the headline shape: 5.60-5.64 ms / step fused against 5.57 unfused (the extra 25 MB store per subnet in the
# epilogue costs what the separate coupling kernel did), so the default keeps the plain convolution + coupling_apply.
FUSE_STORE = os.environ.get("SININN_FUSE_STORE", "0") != "0"
# A channel permutation next to a coupling block whose adjoining half-step runs as a standalone kernel is folded into that
# kernel (coupling_apply_permute / coupling_bwd_unpermute): every channel moves once instead of twice
FOLD_PERM = os.environ.get("SININN_FOLD_PERM", "1") != "0"
# two squeezes ahead of the first coupling block + the change to channels-last (and back) as one kernel
FUSE_SQUEEZE2 = os.environ.get("SININN_FUSE_SQUEEZE2", "1") != "0"

TRACE = None     # debugging aid: set to a list to collect (label, trunk copy) after every executed op


def _trace(label, t):
    if TRACE is not None:
        TRACE.append((label, t.detach().float().cpu().clone()))


class LatentInput:
    """Input of the inverse pass given as its two parts, (lr, z) of lit_wrapper.py:41-42 / 110-111, instead of their
    concatenation: the cat (and, with z=None, the draw z = temp * N(0,1)) happens inside the kernel that brings the
    input into the channels-last trunk.  step_state: device int32 counter that advances the random stream (a CUDA-graph
    replay then draws a fresh z each step).  Not differentiable w.r.t. lr / z (the reference never asks for that)."""

    def __init__(self, lr, z=None, z_dims=None, temp=1.0, seed=0, offset=0, step_state=None):
        if z is None and z_dims is None:
            raise SininnError("LatentInput: give z or z_dims")
        self.lr, self.z = lr, z
        self.z_dims = z.shape[1] if z is not None else int(z_dims)
        self.temp, self.seed, self.offset, self.step_state = float(temp), int(seed), int(offset), step_state

    @property
    def device(self):
        return self.lr.device

    @property
    def is_cuda(self):
        return self.lr.is_cuda

    @property
    def requires_grad(self):
        return False

    @property
    def shape(self):
        b, l, h, w = self.lr.shape
        return torch.Size((b, l + self.z_dims, h, w))

    @property
    def dtype(self):
        return self.lr.dtype

    def dim(self):
        return 4

    def detach(self):
        return self


# ----------------------------------------------------------------------------- packed-weight cache
_pack_cache = {}     # id(param) -> (weakref(param), {(mode, dtype): (version, data_ptr, packed)})


def invalidate_packs():
    """Forget packed weights.  Call after parameters were modified outside torch's version tracking: an update through
    ``param.data`` or through a raw pointer (the fused Adam kernel, a replayed CUDA graph) does not bump ``_version``,
    so the packed bf16 copies the kernels read would silently stay stale."""
    _pack_cache.clear()
    _pack_epoch[0] += 1


def packed(w, mode, dtype):
    """Implicit-GEMM layout of an OIHW weight, cached per parameter OBJECT until it changes (version counter
    or storage).  Keyed on identity, not on the address: the caching allocator re-uses addresses across nets."""
    key = id(w)
    ent = _pack_cache.get(key)
    if ent is None or ent[0]() is not w:
        ent = (weakref.ref(w, lambda _r, k=key: _pack_cache.pop(k, None)), {})
        _pack_cache[key] = ent
    hit = ent[1].get((mode, dtype))
    if hit is not None and hit[0] == w._version and hit[1] == w.data_ptr():
        return hit[2]
    co, ci = w.shape[0], w.shape[1]
    rows, k = (co, ci) if mode % 2 == 0 else (ci, co)      # (mode 4 = fprop with interleaved rows)
    p = K.pack_weight(w, mode, dtype, _round_up(rows, 16), _round_up(k, 16) if mode not in (2, 3) else K.SPLIT_BLOCKS * _round_up(k, 8))
    ent[1][(mode, dtype)] = (w._version, w.data_ptr(), p)
    return p


_pack_epoch = [0]


class PackSet:
    """All conv weights of one plan, re-laid-out for the implicit GEMMs by ONE batched kernel launch whenever any
    of them changed (fprop and dgrad layouts).  Outputs and the device-side job table are allocated once."""

    def __init__(self, weights, dtype, split=False):
        self.weights, self.dtype, self.split = list(weights), dtype, split
        self.out, self.jobs, self.sig, self.ptrs = {}, None, None, None

    def _build(self, dev):
        rows = []
        self.out = {}
        for w in self.weights:
            co, ci, kh, kw = w.shape
            for mode in (0, 1):
                r, k = (co, ci) if mode == 0 else (ci, co)
                rp, kp = _round_up(r, 16), _round_up(k, 16)
                if self.split:                       # pack modes 2 / 3: K = six blocks of round8(k)
                    kp = K.SPLIT_BLOCKS * _round_up(k, 8)
                t = torch.empty(kh * kw, rp, kp, dtype=self.dtype, device=dev)
                self.out[(id(w), mode)] = t
                rows.append([w.data_ptr(), t.data_ptr(), co, ci, kh * kw, mode + (2 if self.split else 0), rp, kp])
            if not self.split and self.dtype == torch.bfloat16 and co % 2 == 0 and co <= 256:
                # interleaved fprop pack (rows s_0, t_0, s_1, t_1, ...) for the fused coupling epilogue
                rp, kp = _round_up(co, 16), _round_up(ci, 16)
                t = torch.empty(kh * kw, rp, kp, dtype=self.dtype, device=dev)
                self.out[(id(w), 4)] = t
                rows.append([w.data_ptr(), t.data_ptr(), co, ci, kh * kw, 4, rp, kp])
        self.jobs = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.ptrs = tuple(w.data_ptr() for w in self.weights)

    def ensure(self):
        if not self.weights:
            return
        ptrs = tuple(w.data_ptr() for w in self.weights)
        if ptrs != self.ptrs:
            self._build(self.weights[0].device)
            self.sig = None
        sig = (tuple(w._version for w in self.weights), _pack_epoch[0])
        if sig != self.sig:
            K.pack_weights_batched(self.jobs, self.jobs.shape[0], self.dtype)
            self.sig = sig

    def get(self, w, mode):
        return self.out[(id(w), mode)]


# ----------------------------------------------------------------------------- trunk state
class Trunk:
    """fp32 channels-last activations [B,h,w,C] (+ gradient of the same shape during backward)
    and a cache of compact bf16 copies of channel ranges (tensor-core operands)."""

    def __init__(self, U, dU=None):
        self.U, self.dU = U, dU
        self.bf = {}

    @property
    def geom(self):
        return tuple(self.U.shape[:3])

    @property
    def C(self):
        return self.U.shape[3]

    @property
    def npix(self):
        s = self.U.shape
        return s[0] * s[1] * s[2]

    def mat(self):
        return self.U.view(self.npix, self.C)

    def dmat(self):
        return self.dU.view(self.npix, self.C)

    def set(self, U, dU=None, bf=None):
        self.U, self.dU = U, dU
        self.bf = bf or {}

    def invalidate(self, c0, c1):
        for key in [k for k in self.bf if not (k[1] <= c0 or k[0] >= c1)]:
            del self.bf[key]

    def split_operand(self, rng):
        """Channel range as a split bf16 operand [npix, 6*round8(L)] (fp32-accurate tensor-core path); cached like the
        plain bf16 copies and dropped by invalidate()."""
        key = (rng[0], rng[1], "split")
        hit = self.bf.get(key)
        if hit is None:
            hit = K.split_bf16(self.mat()[:, rng[0]:rng[1]])
            self.bf[key] = hit
        return hit

    def operand(self, rng, dtype):
        """Channel range as a GEMM operand of `dtype` ([npix, L] view)."""
        c0, c1 = rng
        if dtype == torch.float32:
            return self.mat()[:, c0:c1]
        hit = self.bf.get((c0, c1))
        if hit is None:
            hit = torch.empty(self.npix, _round_up(c1 - c0, 8), dtype=dtype, device=self.U.device)[:, :c1 - c0]
            K.cast_slice(self.mat()[:, c0:c1], hit)
            self.bf[(c0, c1)] = hit
        return hit


class RunCtx:
    def __init__(self, cfg, want_grads=False, packs=None, direct_grad=False, split_packs=None, stash=None):
        self.cfg = cfg
        # "store" mode: the value pass appends every subnet's (output, saved tensors) here, backward pops them
        self.stash = stash
        self.adt = cfg.act_dtype
        self.tc = cfg.tc
        self.split = cfg.split
        self.split_packs = split_packs
        self.grads = {} if want_grads else None
        self.packs = packs
        self.direct_grad = direct_grad      # accumulate weight gradients straight into param.grad (no autograd add)
        self.direct = set()
        self.wstream = None                 # side stream for the parameter-gradient launches (Plan.backward sets it)
        # tensor-core path: weight gradients are collected and launched in groups of up to WGRAD_GROUP problems (the
        # convolutions of a coupling block): fewer pixel splits per problem, see sininn_wgrad_tc_group
        self.pending = []

    def grad_out(self, param):
        """(tensor, accumulate) the weight-gradient kernels should write to for `param`."""
        g = param.grad
        if (self.direct_grad and g is not None and g.dtype == torch.float32 and g.is_contiguous()
                and g.shape == param.shape and g.device == param.device):
            self.direct.add(id(param))
            return g, True
        cur = self.grads.get(id(param))
        if cur is not None:
            return cur, True
        t = torch.empty_like(param, dtype=torch.float32)
        self.grads[id(param)] = t
        return t, False

    def pack(self, w, mode):
        """Implicit-GEMM layout of conv weight w (mode 0 fprop, 1 dgrad) in the activation dtype."""
        if self.packs is not None:
            return self.packs.get(w, mode)
        return packed(w, mode, self.adt)

    def pack_split(self, w, mode):
        """bf16 hi/lo split layout (pack modes 2 / 3) for the fp32-accurate tensor-core path."""
        if self.split_packs is not None:
            return self.split_packs.get(w, mode)
        return packed(w, mode + 2, torch.bfloat16)

    def add_grad(self, param, g):
        if param is None or not param.requires_grad:
            return
        cur = self.grads.get(id(param))
        self.grads[id(param)] = g if cur is None else cur + g


# ----------------------------------------------------------------------------- subnets
WGRAD_GROUP = max(1, min(4, int(os.environ.get("SININN_WGRAD_GROUP", "4"))))
# DenseBlock (IRN): the weight gradients of its five convolutions as one problem over the shared concatenation
MERGE_DENSE_WGRAD = os.environ.get("SININN_MERGE_DENSE_WGRAD", "1") != "0"


def flush_param_grads(ctx):
    """Launch the collected weight/bias gradients as one group (on the side stream when there is one)."""
    jobs, ctx.pending = ctx.pending, []
    if not jobs:
        return
    if ctx.wstream is not None:
        cur = torch.cuda.current_stream()
        ctx.wstream.wait_stream(cur)
        with torch.cuda.stream(ctx.wstream):
            K.wgrad_group(jobs)
        for j in jobs:
            for t in (j[0], j[1]):
                t.record_stream(ctx.wstream)      # keep the operands alive until the side stream has consumed them
        return
    K.wgrad_group(jobs)


def _param_grads_split(ctx, conv, xs, dys, cinp, coutp, geom, taps):
    """fp32-accurate weight/bias gradient on the tensor cores: the weight-gradient kernel walks the pixels once per
    product term of the three-term expansions x = h+m+l, dy = dh+dm+dl (h*dh, h*dm, m*dh, m*dm, h*dl, l*dh: everything down
    to second order), reading the matching channel blocks of the split operands; the bias gradient sums dh, dm, dl."""
    wants_w = conv.weight.requires_grad
    wants_b = conv.bias is not None and conv.bias.requires_grad
    if not (wants_w or wants_b):
        return
    g, acc = ctx.grad_out(conv.weight) if wants_w else (torch.empty_like(conv.weight), False)
    gb, accb = ctx.grad_out(conv.bias) if wants_b else (None, False)
    H_, M_, L_ = 0, 1, 3                      # block index of each term in [h | m | h | l (| m | h)]
    xo = [H_ * cinp, H_ * cinp, M_ * cinp, M_ * cinp, H_ * cinp, L_ * cinp]
    yo = [H_ * coutp, M_ * coutp, H_ * coutp, M_ * coutp, L_ * coutp, H_ * coutp]
    terms = (xo, yo, 0b010011)                # dh (pair 0), dm (pair 1), dl (pair 4) enter the bias sum once each
    ctx.pending.append((xs[:, :conv.in_channels], dys[:, :conv.out_channels], geom, taps, g, acc, gb, accb, terms))
    if len(ctx.pending) >= WGRAD_GROUP:
        flush_param_grads(ctx)


def _param_grads(ctx, conv, x, dy, geom, taps):
    """Weight and bias gradients of one conv from its input x and output gradient dy.  On the tensor-core path the
    bias gradient (column sums of dy) is produced by the weight-gradient launches themselves; otherwise by colsum."""
    wants_w = conv.weight.requires_grad
    wants_b = conv.bias is not None and conv.bias.requires_grad
    if ctx.tc and wants_w and WGRAD_GROUP > 1 and x.stride(0) % 8 == 0 and dy.stride(0) % 8 == 0:
        g, acc = ctx.grad_out(conv.weight)
        gb, accb = ctx.grad_out(conv.bias) if wants_b else (None, False)
        ctx.pending.append((x, dy, geom, taps, g, acc, gb, accb))
        if len(ctx.pending) >= WGRAD_GROUP:
            flush_param_grads(ctx)
        return
    if ctx.wstream is not None and (wants_w or wants_b):
        # parameter gradients are leaves of the backward pass: nothing downstream reads them before the pass ends,
        # so they go to a side stream and fill the gaps of the data-gradient chain (joined in Plan.backward)
        cur = torch.cuda.current_stream()
        ctx.wstream.wait_stream(cur)
        with torch.cuda.stream(ctx.wstream):
            _param_grads_here(ctx, conv, x, dy, geom, taps, wants_w, wants_b)
        for t in (x, dy):
            t.record_stream(ctx.wstream)          # keep the operands alive until the side stream has consumed them
        return
    _param_grads_here(ctx, conv, x, dy, geom, taps, wants_w, wants_b)


def _param_grads_here(ctx, conv, x, dy, geom, taps, wants_w, wants_b):
    if wants_w:
        g, acc = ctx.grad_out(conv.weight)
        if wants_b and ctx.tc:
            gb, accb = ctx.grad_out(conv.bias)
            K.wgrad(x, dy, geom, taps, g, accumulate=acc, tensor_core=True, dbias=gb, dbias_accumulate=accb)
            return
        K.wgrad(x, dy, geom, taps, g, accumulate=acc, tensor_core=ctx.tc)
    if wants_b:
        gb, accb = ctx.grad_out(conv.bias)
        K.colsum(dy, gb, accumulate=accb)


def _is_conv(m, k=None):
    return (isinstance(m, nn.Conv2d) and m.kernel_size[0] == m.kernel_size[1] and m.kernel_size[0] in (1, 3)
            and m.stride == (1, 1) and m.dilation == (1, 1) and m.groups == 1
            and m.padding == (m.kernel_size[0] // 2,) * 2 and m.padding_mode == "zeros"
            and (k is None or m.kernel_size[0] == k))


class ConvSubnet:
    """conv(k) -> ReLU -> conv(k): subnet_conv / subnet_conv_1x1 (archs.py:11-17)."""

    def __init__(self, seq):
        mods = list(seq.children()) if isinstance(seq, nn.Sequential) else []
        if not (len(mods) == 3 and _is_conv(mods[0]) and isinstance(mods[1], nn.ReLU) and _is_conv(mods[2], mods[0].kernel_size[0])):
            raise SininnError("unsupported coupling subnet: expected nn.Sequential(Conv2d(k), ReLU, Conv2d(k)) with "
                              f"k in (1,3), stride 1, 'same' zero padding; got {seq}")
        self.c1, self.c2 = mods[0], mods[2]
        self.taps = self.c1.kernel_size[0] ** 2
        self.cin, self.hidden, self.cout = self.c1.in_channels, self.c1.out_channels, self.c2.out_channels
        self._fused_bwd = None

    def parameters(self):
        return [p for p in (self.c1.weight, self.c1.bias, self.c2.weight, self.c2.bias) if p is not None]

    def stash_bytes_per_pixel(self, cfg):
        if cfg.split:
            return 12 * _round_up(self.cin, 8) + 8 * self.hidden + self.hidden // 8 + 4 * self.cout
        e = 2 if cfg.tc else 4
        if cfg.tc and self._bwd_fusable():
            return e * self.cin + 4 * self.cout           # the hidden activation is re-evaluated on chip
        return e * (self.cin + self.hidden) + self.hidden // 8 + 4 * self.cout

    def _bwd_fusable(self):
        """Shape / parameter test for the fused backward kernel (the caller checks that the tensor-core path is on).
        Decided when the value pass runs: a subnet whose backward is fused keeps no hidden activation to fall back on."""
        if self._fused_bwd is None:
            self._fused_bwd = bool(
                FUSE_1X1_BWD and FUSE_1X1 and self.taps == 1 and self.c1.bias is not None and self.c2.bias is not None
                and self.cin % 8 == 0 and self.cout % 8 == 0
                and K.subnet1x1_supported(self.cin, self.hidden, self.cout)
                and K.subnet1x1_bwd_supported(self.cin, self.hidden, self.cout))
        return self._fused_bwd and all(q.requires_grad for q in self.parameters())

    def fused_bwd(self, ctx):
        return bool(ctx.tc and self._bwd_fusable())

    def can_fuse_coupling(self, ctx, L, backward):
        mask = FUSE_COUPLING if self.taps == 9 else FUSE_COUPLING_1X1
        on = (mask & 1) if not backward else ((mask & 2) or ((mask & 4) and self.cout <= 64))
        if not (ctx.tc and self.cout == 2 * L and L % 8 == 0 and self.cout <= 256 and bool(on) and not getattr(self, "_no_fuse", False)):
            return False
        return self.taps == 9 or (FUSE_1X1 and K.subnet1x1_supported(self.cin, self.hidden, self.cout))

    def fwd_coupled(self, ctx, tr, src, u, clamp, rev, want_bf, du=None, store=False):
        """Subnet forward with the GLOW half-step (du None) or its backward (du given) applied to u in the second
        convolution's epilogue: the subnet output [s | t] never leaves the SM -- unless `store` asks for the copy a
        backward pass in "store" mode reads.  Returns (bf16 copy of the new u or None, saved tensors for bwd() or None,
        da (backward) / a (store) or None); None altogether if the kernel does not take the shape."""
        x = tr.operand(src, ctx.adt)
        dev = x.device
        L = u.shape[1]
        keep = du is not None or store
        bf = torch.empty(tr.npix, L, dtype=torch.bfloat16, device=dev) if want_bf else None
        da = torch.empty(tr.npix, 2 * L, dtype=torch.bfloat16, device=dev) if du is not None else None
        cpl = dict(mode=2 if du is not None else 1, u=u, clamp=clamp, inverse=rev, bf16=bf, du=du, da=da)
        if store:
            da = cpl["a"] = torch.empty(tr.npix, 2 * L, dtype=torch.float32, device=dev)
        if self.taps == 1:
            if x.stride(0) % 8 != 0:
                return None
            # fused 1x1 subnet: the hidden tile stays in shared memory AND the subnet output stays in registers
            keep_h = keep and not self.fused_bwd(ctx)       # (the fused backward kernel re-evaluates h from x on chip)
            h = torch.empty(tr.npix, self.hidden, dtype=ctx.adt, device=dev) if keep_h else None
            bits = torch.empty(tr.npix, self.hidden // 32, dtype=torch.int32, device=dev) if keep_h else None
            K.subnet1x1_fwd(x, ctx.pack(self.c1.weight, 0), self.c1.bias, ctx.pack(self.c2.weight, 4), self.c2.bias, None,
                            h_out=h, bits_out=bits, coupling=cpl)
            return bf, ((x, h, bits) if keep else None), da
        h = torch.empty(tr.npix, self.hidden, dtype=ctx.adt, device=dev)
        bits = torch.empty(tr.npix, (self.hidden + 31) // 32, dtype=torch.int32, device=dev) if keep else None
        K.conv(x, ctx.pack(self.c1.weight, 0), tr.geom, self.hidden, h, bias=self.c1.bias, act=ACT_RELU, tensor_core=True, bits_out=bits)
        try:
            K.conv(h, ctx.pack(self.c2.weight, 4), tr.geom, self.cout, None, bias=self.c2.bias, tensor_core=True, coupling=cpl)
        except SininnError as e:
            if "code -3" not in str(e):
                raise
            self._no_fuse = True                     # shape not taken by the CTA-pair kernel: unfused path from now on
            return None
        return bf, ((x, h, bits) if keep else None), da

    def _fwd_split(self, ctx, tr, src, keep):
        """fp32-accurate forward on the tensor cores: both convolutions on split operands, hidden activation in fp32.
        The first convolution takes the full six-block product (its output decides the ReLUs); everything downstream
        of the ReLU -- second convolution, both data gradients -- takes the four-block (two-term-weight) product."""
        dev = tr.U.device
        xs = tr.split_operand(src)
        h = torch.empty(tr.npix, self.hidden, dtype=torch.float32, device=dev)
        bits = torch.empty(tr.npix, (self.hidden + 31) // 32, dtype=torch.int32, device=dev) if keep else None
        K.conv(xs, ctx.pack_split(self.c1.weight, 0), tr.geom, self.hidden, h, bias=self.c1.bias, act=ACT_RELU, tensor_core=True,
               bits_out=bits)
        hs = K.split_bf16(h, blocks=4)
        a = torch.empty(tr.npix, self.cout, dtype=torch.float32, device=dev)
        K.conv(hs, ctx.pack_split(self.c2.weight, 0), tr.geom, self.cout, a, bias=self.c2.bias, tensor_core=True)
        return a, (xs, hs, bits)

    def _bwd_split(self, ctx, tr, saved, da, dsrc):
        xs, hs, bits = saved
        dev = xs.device
        das = K.split_bf16(da, blocks=4)
        dh = torch.empty(tr.npix, self.hidden, dtype=torch.float32, device=dev)
        K.conv(das, ctx.pack_split(self.c2.weight, 1), tr.geom, self.hidden, dh, mask_bits=bits, tensor_core=True)
        _param_grads_split(ctx, self.c2, hs, das, _round_up(self.hidden, 8), _round_up(self.cout, 8), tr.geom, self.taps)
        dhs = K.split_bf16(dh, blocks=4)
        K.conv(dhs, ctx.pack_split(self.c1.weight, 1), tr.geom, self.cin, dsrc, accumulate=True, tensor_core=True)
        _param_grads_split(ctx, self.c1, xs, dhs, _round_up(self.cin, 8), _round_up(self.hidden, 8), tr.geom, self.taps)

    def fwd(self, ctx, tr, src, keep=False):
        if ctx.split:
            return self._fwd_split(ctx, tr, src, keep)
        x = tr.operand(src, ctx.adt)
        dev = x.device
        if (ctx.tc and self.taps == 1 and FUSE_1X1 and K.subnet1x1_supported(self.cin, self.hidden, self.cout)
                and x.stride(0) % 8 == 0):
            # one launch for the whole subnet: the hidden activation stays in shared memory (it is only written
            # out, with its ReLU sign bits, when the backward kernels need it)
            a = torch.empty(tr.npix, self.cout, dtype=torch.float32, device=dev)
            keep_h = keep and not self.fused_bwd(ctx)       # (the fused backward kernel re-evaluates h from x on chip)
            h = torch.empty(tr.npix, self.hidden, dtype=ctx.adt, device=dev) if keep_h else None
            bits = torch.empty(tr.npix, self.hidden // 32, dtype=torch.int32, device=dev) if keep_h else None
            K.subnet1x1_fwd(x, ctx.pack(self.c1.weight, 0), self.c1.bias, ctx.pack(self.c2.weight, 0), self.c2.bias, a,
                            h_out=h, bits_out=bits)
            return a, (x, h, bits)
        h = torch.empty(tr.npix, self.hidden, dtype=ctx.adt, device=dev)
        # tensor-core path: the backward pass reads the ReLU mask as 1 bit / element instead of re-reading h
        bits = (torch.empty(tr.npix, (self.hidden + 31) // 32, dtype=torch.int32, device=dev)
                if (keep and ctx.tc) else None)
        K.conv(x, ctx.pack(self.c1.weight, 0), tr.geom, self.hidden, h, bias=self.c1.bias, act=ACT_RELU,
               tensor_core=ctx.tc, bits_out=bits)
        a = torch.empty(tr.npix, self.cout, dtype=torch.float32, device=dev)
        K.conv(h, ctx.pack(self.c2.weight, 0), tr.geom, self.cout, a, bias=self.c2.bias, tensor_core=ctx.tc)
        if ctx.stash is not None and x.dtype == torch.float32:
            x = x.clone()                            # fp32 operands are views of the live trunk; a stored one must be private
        return a, (x, h, bits)

    def bwd(self, ctx, tr, saved, da, dsrc):
        """da: dL/d(output) [npix, cout] in the activation dtype; accumulates dL/d(input) into dsrc (fp32 view)."""
        if ctx.split:
            return self._bwd_split(ctx, tr, saved, da, dsrc)
        x, h, bits = saved
        dev = x.device
        if h is None:
            # 1x1 subnet, fused backward: h re-evaluated on chip, dsrc += W1^T dh and the four parameter gradients in one
            # kernel (+ one fixed-order reduction); neither h nor dh exists in memory
            K.subnet1x1_bwd(x, da, ctx.pack(self.c1.weight, 0), self.c1.bias, ctx.pack(self.c2.weight, 1), ctx.pack(self.c1.weight, 1),
                            dsrc, ctx.grad_out(self.c1.weight) + ctx.grad_out(self.c1.bias),
                            ctx.grad_out(self.c2.weight) + ctx.grad_out(self.c2.bias))
            return
        dh = torch.empty_like(h)
        if (ctx.tc and self.taps == 1 and FUSE_1X1 and bits is not None and da.stride(0) % 8 == 0
                and K.subnet1x1_supported(self.cout, self.hidden, self.cin) and dsrc.stride(0) % 4 == 0
                and dsrc.data_ptr() % 16 == 0):
            # both data gradients in ONE launch of the fused pipeline: dh = mask * (W2^T da) stays in shared memory
            # for dsrc += W1^T dh and is stored once for the weight gradient of conv1
            K.subnet1x1_fwd(da, ctx.pack(self.c2.weight, 1), None, ctx.pack(self.c1.weight, 1), None, dsrc,
                            h_out=dh, mask_bits=bits, accumulate=True)
            _param_grads(ctx, self.c2, h, da, tr.geom, self.taps)
            _param_grads(ctx, self.c1, x, dh, tr.geom, self.taps)
            return
        if bits is not None:
            K.conv(da, ctx.pack(self.c2.weight, 1), tr.geom, self.hidden, dh, mask_bits=bits, tensor_core=True)
        else:
            K.conv(da, ctx.pack(self.c2.weight, 1), tr.geom, self.hidden, dh, mask=h, mask_act=ACT_RELU,
                   tensor_core=ctx.tc)
        _param_grads(ctx, self.c2, h, da, tr.geom, self.taps)
        K.conv(dh, ctx.pack(self.c1.weight, 1), tr.geom, self.cin, dsrc, accumulate=True, tensor_core=ctx.tc)
        _param_grads(ctx, self.c1, x, dh, tr.geom, self.taps)


class DenseSubnet:
    """DenseBlock (archs.py:74-95): five 3x3 convs over a growing channel concatenation, LeakyReLU(0.2)
    after the first four.  The concatenation is one channels-last buffer; conv j reads its first
    cin+32(j-1) channels and writes its 32 outputs right behind them, so torch.cat costs nothing."""

    SLOPE = 0.2

    def __init__(self, block):
        self.convs = [block.conv1, block.conv2, block.conv3, block.conv4, block.conv5]
        for c in self.convs:
            if not _is_conv(c, 3):
                raise SininnError(f"unsupported DenseBlock conv {c}")
        self.cin = self.convs[0].in_channels
        self.gc = self.convs[0].out_channels
        self.cout = self.convs[4].out_channels
        self.ctot = self.cin + 4 * self.gc
        self._merged_ok = None

    def parameters(self):
        out = []
        for c in self.convs:
            out += [p for p in (c.weight, c.bias) if p is not None]
        return out

    def stash_bytes_per_pixel(self, cfg):
        return (2 if cfg.tc else 4) * _round_up(self.ctot, 8) + 4 * self.cout

    def fwd(self, ctx, tr, src, keep=False):
        dev = tr.U.device
        cat = torch.empty(tr.npix, _round_up(self.ctot, 8), dtype=ctx.adt, device=dev)
        K.cast_slice(tr.mat()[:, src[0]:src[1]], cat[:, :self.cin])
        for j in range(4):
            lo = self.cin + self.gc * j
            K.conv(cat[:, :lo], ctx.pack(self.convs[j].weight, 0), tr.geom, self.gc, cat[:, lo:lo + self.gc],
                   bias=self.convs[j].bias, act=ACT_LRELU, slope=self.SLOPE, tensor_core=ctx.tc)
        out = torch.empty(tr.npix, self.cout, dtype=torch.float32, device=dev)
        K.conv(cat[:, :self.ctot], ctx.pack(self.convs[4].weight, 0), tr.geom, self.cout, out,
               bias=self.convs[4].bias, tensor_core=ctx.tc)
        return out, (cat,)

    def _merge_wgrads(self, ctx, cat):
        """One weight-gradient problem for all five convolutions (they read prefixes of the same concatenation)?"""
        if not (MERGE_DENSE_WGRAD and ctx.tc and WGRAD_GROUP > 1 and len(self.convs) <= 8):
            return False
        if any(c.bias is None or not c.weight.requires_grad or not c.bias.requires_grad for c in self.convs):
            return False
        if self._merged_ok is None:
            probe = torch.empty(8, _round_up(4 * self.gc + self.cout, 8), dtype=cat.dtype, device=cat.device)
            self._merged_ok = bool(K.wgrad_merged_supported(cat[:8, :self.ctot], probe[:, :4 * self.gc + self.cout], (1, 1, 8), 9))
        return self._merged_ok

    def bwd(self, ctx, tr, saved, dout, dsrc):
        (cat,) = saved
        dev = cat.device
        # gradient of the concatenation buffer accumulates in fp32; each conv's upstream gradient is
        # cast to the operand dtype (compact, aligned) after the LeakyReLU derivative is applied
        dcat = torch.empty(tr.npix, self.ctot, dtype=torch.float32, device=dev)
        c5 = self.convs[4]
        merged = self._merge_wgrads(ctx, cat)
        if merged:
            # the five output gradients side by side: [g1 | g2 | g3 | g4 | dout], the "dy" of ONE weight-gradient problem
            # over x = the concatenation (launched when all of them exist) instead of five problems with N = 32
            gall = torch.empty(tr.npix, _round_up(4 * self.gc + self.cout, 8), dtype=ctx.adt, device=dev)
            gall[:, 4 * self.gc:4 * self.gc + self.cout].copy_(dout)
        K.conv(dout, ctx.pack(c5.weight, 1), tr.geom, self.ctot, dcat, tensor_core=ctx.tc)
        if not merged:
            _param_grads(ctx, c5, cat[:, :self.ctot], dout, tr.geom, 9)
        for j in (3, 2, 1, 0):
            lo = self.cin + self.gc * j
            g = gall[:, self.gc * j:self.gc * (j + 1)] if merged else torch.empty(tr.npix, self.gc, dtype=ctx.adt, device=dev)
            K.act_bwd(dcat[:, lo:lo + self.gc], cat[:, lo:lo + self.gc], g, ACT_LRELU, self.SLOPE)
            cj = self.convs[j]
            K.conv(g, ctx.pack(cj.weight, 1), tr.geom, lo, dcat[:, :lo], accumulate=True, tensor_core=ctx.tc)
            if not merged:
                _param_grads(ctx, cj, cat[:, :lo], g, tr.geom, 9)
        K.axpy_slice(dsrc, dcat[:, :self.cin], 1.0)
        if merged:
            segs = []
            for j, cj in enumerate(self.convs):
                gw, accw = ctx.grad_out(cj.weight)
                gb, accb = ctx.grad_out(cj.bias)
                rows = self.gc if j < 4 else self.cout
                segs.append((self.gc * j, rows, self.cin + self.gc * j, gw, accw, gb, accb))
            x, dy = cat[:, :self.ctot], gall[:, :4 * self.gc + self.cout]
            if ctx.wstream is not None:       # a leaf of the backward pass: next to the data-gradient chain (see _param_grads)
                ctx.wstream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(ctx.wstream):
                    K.wgrad_merged(x, dy, tr.geom, 9, segs)
                cat.record_stream(ctx.wstream)
                gall.record_stream(ctx.wstream)
            else:
                K.wgrad_merged(x, dy, tr.geom, 9, segs)


# ----------------------------------------------------------------------------- plan ops
class ResampleOp:
    """mode 0: FrEIA IRevNetDownsampling (archs.py:28-31,35-38); mode 1: HaarDownsampling (archs.py:162-199)."""
    kind = "resample"

    def __init__(self, mode):
        self.mode = mode

    def fwd_scale(self):
        return 0.25 if self.mode == 1 else 1.0

    # value maps: forward scale s_f, inverse scale 1.  Gradient maps are the transposes:
    #   grad of forward  = inverse map * s_f ; grad of inverse = forward map / s_f  (H H^T = 4 I for Haar)
    def apply_nchw(self, x, rev, grad=False):
        if not grad:
            return K.resample_nchw(x, self.mode, rev, 1.0 if rev else self.fwd_scale())
        if not rev:   # gradient of the forward map: [B,4C,h,w] -> [B,C,2h,2w]
            return K.resample_nchw(x, self.mode, 1, self.fwd_scale())
        return K.resample_nchw(x, self.mode, 0, 1.0 if self.mode == 0 else 1.0)

    def apply_nhwc(self, x, rev, grad=False):
        if not grad:
            return K.resample_nhwc(x, self.mode, rev, 1.0 if rev else self.fwd_scale())
        if not rev:
            return K.resample_nhwc(x, self.mode, 1, self.fwd_scale())
        return K.resample_nhwc(x, self.mode, 0, 1.0)

    def parameters(self):
        return []


class PermOp:
    """FrEIA PermuteRandom (archs.py:65-68): fwd x[:, perm], rev x[:, perm_inv]."""
    kind = "perm"

    def __init__(self, perm):
        self.perm_cpu = torch.as_tensor(perm, dtype=torch.int32).clone()
        inv = torch.empty_like(self.perm_cpu)
        inv[self.perm_cpu.long()] = torch.arange(len(self.perm_cpu), dtype=torch.int32)
        self.inv_cpu = inv
        self._dev = {}

    def maps(self, device):
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = (self.perm_cpu.to(device), self.inv_cpu.to(device))
        return self._dev[key]

    def gather_map(self, device, rev, grad=False):
        """Index map m with out[:, i] = in[:, m[i]] for the value pass (or its gradient pass)."""
        perm, inv = self.maps(device)
        use_inv = bool(rev) != bool(grad)     # gradient of a gather with perm is a gather with perm_inv
        return inv if use_inv else perm

    def parameters(self):
        return []


class LinearOp:
    """Fixed invertible 1x1 convolution (FrEIA Fixed1x1Conv, offered at archs.py:40-50): y[:, o] = sum_i M[i][o] x[:, i]
    per pixel, reverse with M^-1 (inverted in fp64 on the host).  Always evaluated by the fp32 CUDA-core GEMM kernel,
    whatever the subnet precision: the exactness of the round trip (1e-5) rests on it and C <= 192 makes it cheap.
    PermuteRandom is the special case of a permutation matrix."""
    kind = "linear"

    def __init__(self, M):
        M = torch.as_tensor(M, dtype=torch.float64)
        if M.dim() != 2 or M.shape[0] != M.shape[1]:
            raise SininnError(f"Fixed1x1Conv needs a square matrix, got {tuple(M.shape)}")
        sign, logabs = torch.linalg.slogdet(M)
        if sign == 0:
            raise SininnError("Fixed1x1Conv: the matrix is singular")
        self.C = M.shape[0]
        self.logdet = float(logabs)                       # log |det M|, per pixel
        W = M.t().contiguous()                            # y = W x
        Wi = torch.linalg.inv(W)
        # matrices applied to a pixel's channel vector, by (reverse direction?, gradient pass?)
        self.mats = {(False, False): W, (True, False): Wi, (False, True): W.t().contiguous(), (True, True): Wi.t().contiguous()}
        self._packs = {}

    def pack(self, device, rev, grad=False):
        key = (device.type, device.index, bool(rev), bool(grad))
        if key not in self._packs:
            w = self.mats[(bool(rev), bool(grad))].to(torch.float32).reshape(self.C, self.C, 1, 1).to(device)
            self._packs[key] = K.pack_weight(w, 0, torch.float32, _round_up(self.C, 16), _round_up(self.C, 16))
        return self._packs[key]

    def pack_split(self, device, rev, grad=False):
        key = (device.type, device.index, bool(rev), bool(grad), "split")
        if key not in self._packs:
            w = self.mats[(bool(rev), bool(grad))].to(torch.float32).reshape(self.C, self.C, 1, 1).to(device)
            self._packs[key] = K.pack_weight(w, 2, torch.bfloat16, _round_up(self.C, 16), K.SPLIT_BLOCKS * _round_up(self.C, 8))
        return self._packs[key]

    def apply(self, U, rev, grad=False, split=False):
        """U: channels-last [B,h,w,C] fp32 -> new tensor of the same shape.  split: on the tensor cores with bf16
        hi/mid/lo operands (north-star item 3: the invertible 1x1 convolution as a tcgen05 GEMM at fp32 accuracy);
        otherwise the CUDA-core fp32 GEMM."""
        out = torch.empty_like(U)
        npix = U.numel() // self.C
        if split:
            K.conv(K.split_bf16(U.view(npix, self.C)), self.pack_split(U.device, rev, grad), tuple(U.shape[:3]), self.C,
                   out.view(npix, self.C), tensor_core=True)
        else:
            K.conv(U.view(npix, self.C), self.pack(U.device, rev, grad), tuple(U.shape[:3]), self.C, out.view(npix, self.C),
                   tensor_core=False)
        return out

    def parameters(self):
        return []


class ActNormOp:
    """FrEIA ActNorm (offered, commented out, at archs.py:40-44): y = x * exp(scale_c) + bias_c, reverse
    (y - bias_c) / exp(scale_c); log|det J| = H*W*sum(scale).  scale / bias are trainable [1, C, 1, 1] parameters; with
    `module.init_on_next_batch` the first batch sets them so that the output has zero mean / unit std per channel."""
    kind = "actnorm"

    def __init__(self, module):
        self.m = module

    def parameters(self):
        return [self.m.scale, self.m.bias]

    def _maybe_init(self, tr):
        m = self.m
        if getattr(m, "init_on_next_batch", False):
            with torch.no_grad():                       # one-time statistics of the first batch (plain torch reductions)
                flat = tr.mat()
                ls = torch.log(1.0 / flat.std(dim=0))
                m.scale.data.view(-1).copy_(ls)
                m.bias.data.view(-1).copy_(-(flat * ls.exp()).mean(dim=0))
            m.init_on_next_batch = False

    def apply(self, tr, rev):
        self._maybe_init(tr)
        tr.bf = {}
        K.channel_affine(tr.U, self.m.scale.detach().view(-1), self.m.bias.detach().view(-1), rev)

    def backward(self, ctx, tr, rev):
        m = self.m
        gs, acc_s = ctx.grad_out(m.scale)
        gb, acc_b = ctx.grad_out(m.bias)
        if acc_s != acc_b:                              # (both come from the same arena or both are fresh)
            raise SininnError("ActNorm: scale and bias gradients must be accumulated the same way")
        tr.bf = {}
        K.channel_affine_bwd(tr.U, tr.dU, m.scale.detach().view(-1), m.bias.detach().view(-1), rev, gs.view(-1), gb.view(-1), acc_s)


class HalfStep:
    """dst <- affine(dst; nets(src)).  kind: 'glow' (one net, output = [s | t]), 'irn_affine' (s from
    nets[0], t from nets[1]), 'irn_add' (dst <- dst + sign * nets[0](src))."""

    def __init__(self, kind, nets, src, dst, clamp):
        self.kind, self.nets, self.src, self.dst, self.clamp = kind, nets, src, dst, clamp


class CouplingOp:
    kind = "coupling"

    def __init__(self, steps_fwd, channels):
        self.steps = steps_fwd          # execution order for rev=False; rev=True runs them reversed + inverted
        self.channels = channels

    def parameters(self):
        out = []
        seen = set()
        for st in self.steps:
            for n in st.nets:
                if id(n) not in seen:
                    seen.add(id(n))
                    out += n.parameters()
        return out

    def first_src(self, rev):
        return (self.steps[-1] if rev else self.steps[0]).src

    def stash_bytes_per_pixel(self, cfg):
        return sum(n.stash_bytes_per_pixel(cfg) for st in self.steps for n in st.nets)

    def first_step_backward(self, rev):
        """The half-step whose subnet the backward pass of a value pass `rev` evaluates first."""
        return self.steps[0] if rev else self.steps[-1]

    # ---- value pass
    def _uses_epilogue(self, ctx, st, L, keep, logdet):
        """Value pass: will this GLOW half-step run inside the second convolution's epilogue?"""
        return logdet is None and (FUSE_STORE or not keep) and st.nets[0].can_fuse_coupling(ctx, L, False)

    def folds_perm_value(self, ctx, tr, rev, logdet=None):
        """Can the permutation that follows this block be folded into its last half-step (value pass)?"""
        st = (self.steps[::-1] if rev else self.steps)[-1]
        L = st.dst[1] - st.dst[0]
        return (FOLD_PERM and st.kind == "glow" and tr.C % 4 == 0
                and not self._uses_epilogue(ctx, st, L, ctx.stash is not None, logdet))

    def folds_perm_backward(self, ctx, tr, rev):
        """Can the undo of the permutation that follows this block be folded into the backward of its last half-step?  Only
        with stored subnet outputs: re-evaluating the subnet needs the un-permuted trunk first."""
        st = (self.steps[::-1] if rev else self.steps)[-1]
        L = st.dst[1] - st.dst[0]
        return (FOLD_PERM and ctx.stash is not None and st.kind == "glow" and tr.C % 4 == 0 and L % 4 == 0 and st.dst[0] % 4 == 0)

    def run(self, ctx, tr, rev, logdet=None, post_perm=None):
        """logdet (optional fp32 [B] tensor): accumulates the block's log|det J| per sample (FrEIA's last_jac).
        post_perm = (gather map, bf16 hint range or None): the permutation that follows the block, to be applied by the last
        half-step's kernel (the caller has checked folds_perm_value)."""
        steps = self.steps[::-1] if rev else self.steps
        B = tr.U.shape[0]
        for i, st in enumerate(steps):
            L = st.dst[1] - st.dst[0]
            nxt = steps[i + 1].src if i + 1 < len(steps) else None
            want_bf = ctx.adt == torch.bfloat16 and nxt == st.dst and L % 8 == 0
            u = tr.mat()[:, st.dst[0]:st.dst[1]]
            keep = ctx.stash is not None
            if st.kind == "glow" and post_perm is not None and i == len(steps) - 1:
                a, saved = st.nets[0].fwd(ctx, tr, st.src, keep=keep)
                if keep:
                    ctx.stash.append((a, saved))
                if logdet is not None:
                    K.logscale_sum(a[:, :L], B, GLOW, st.clamp, -1.0 if rev else 1.0, logdet, True)
                cmap, hint = post_perm
                U, bfh = K.coupling_apply_permute(tr.U, cmap, st.dst, a[:, :L], a[:, L:], GLOW, st.clamp, rev, hint, fast=ctx.tc)
                tr.set(U, None, {hint: bfh} if bfh is not None else None)
                _trace(f"half:{st.kind}:{st.src}->{st.dst}+perm", tr.U)
                return
            if st.kind == "glow":
                fused = None
                if self._uses_epilogue(ctx, st, L, keep, logdet):
                    tr.invalidate(*st.dst)
                    fused = st.nets[0].fwd_coupled(ctx, tr, st.src, u, st.clamp, rev, want_bf, store=keep)
                if fused is not None:
                    bf = fused[0]
                    if keep:
                        ctx.stash.append((fused[2], fused[1]))
                else:
                    a, saved = st.nets[0].fwd(ctx, tr, st.src, keep=keep)
                    if keep:
                        ctx.stash.append((a, saved))
                    tr.invalidate(*st.dst)
                    if logdet is not None:
                        K.logscale_sum(a[:, :L], B, GLOW, st.clamp, -1.0 if rev else 1.0, logdet, True)
                    bf = K.coupling_apply(u, a[:, :L], a[:, L:], GLOW, st.clamp, rev, want_bf, fast=ctx.tc)
            elif st.kind == "irn_affine":
                s, saved_s = st.nets[0].fwd(ctx, tr, st.src, keep=keep)
                t, saved_t = st.nets[1].fwd(ctx, tr, st.src, keep=keep)
                if keep:
                    ctx.stash.append((s, saved_s, t, saved_t))
                tr.invalidate(*st.dst)
                if logdet is not None:
                    K.logscale_sum(s, B, IRN, st.clamp, -1.0 if rev else 1.0, logdet, True)
                bf = K.coupling_apply(u, s, t, IRN, st.clamp, rev, want_bf, fast=ctx.tc)
            else:
                f, saved = st.nets[0].fwd(ctx, tr, st.src, keep=keep)
                if keep:
                    ctx.stash.append((f, saved))
                tr.invalidate(*st.dst)
                K.axpy_slice(u, f, -1.0 if rev else 1.0)
                bf = None
            if bf is not None:
                tr.bf[st.dst] = bf
            _trace(f"half:{st.kind}:{st.src}->{st.dst}", tr.U)

    # ---- backward from the output: restores the block input in tr.U and turns tr.dU into dL/d(input)
    def backward(self, ctx, tr, rev, pre_perm=None):
        """pre_perm: gather map that undoes the permutation executed after this block; tr still holds the permuted trunk and
        gradient, and the first half-step's kernel un-permutes them (the caller has checked folds_perm_backward)."""
        executed = self.steps[::-1] if rev else self.steps
        undo = executed[::-1]
        for i, st in enumerate(undo):
            L = st.dst[1] - st.dst[0]
            nxt = undo[i + 1].src if i + 1 < len(undo) else None
            stored = ctx.stash.pop() if ctx.stash is not None else None
            if i == 0 and pre_perm is not None:
                a, saved = stored
                da = torch.empty(tr.npix, 2 * L, dtype=ctx.adt, device=tr.U.device)
                U, dU = K.coupling_bwd_unpermute(tr.U, tr.dU, pre_perm, st.dst, a[:, :L], a[:, L:], GLOW, st.clamp, rev,
                                                 da[:, :L], da[:, L:], fast=ctx.tc)
                tr.set(U, dU)
                st.nets[0].bwd(ctx, tr, saved, da, tr.dmat()[:, st.src[0]:st.src[1]])
                continue
            # with stored subnet internals nothing in backward reads a bf16 copy of the restored trunk
            want_bf = ctx.adt == torch.bfloat16 and nxt == st.dst and L % 8 == 0 and stored is None
            u = tr.mat()[:, st.dst[0]:st.dst[1]]
            du = tr.dmat()[:, st.dst[0]:st.dst[1]]
            dsrc = tr.dmat()[:, st.src[0]:st.src[1]]
            dev = u.device
            if st.kind == "glow":
                fused = None
                if stored is None and st.nets[0].can_fuse_coupling(ctx, L, True):
                    tr.invalidate(*st.dst)
                    fused = st.nets[0].fwd_coupled(ctx, tr, st.src, u, st.clamp, rev, want_bf, du=du)
                if fused is not None:
                    bf, saved, da = fused
                else:
                    a, saved = stored if stored is not None else st.nets[0].fwd(ctx, tr, st.src, keep=True)
                    da = torch.empty(tr.npix, 2 * L, dtype=ctx.adt, device=dev)
                    tr.invalidate(*st.dst)
                    bf = K.coupling_bwd(u, du, a[:, :L], a[:, L:], GLOW, st.clamp, rev, da[:, :L], da[:, L:], want_bf, fast=ctx.tc)
                st.nets[0].bwd(ctx, tr, saved, da, dsrc)
            elif st.kind == "irn_affine":
                if stored is not None:
                    s, saved_s, t, saved_t = stored
                else:
                    s, saved_s = st.nets[0].fwd(ctx, tr, st.src, keep=True)
                    t, saved_t = st.nets[1].fwd(ctx, tr, st.src, keep=True)
                Lp = _round_up(L, 8)
                ds = torch.empty(tr.npix, Lp, dtype=ctx.adt, device=dev)[:, :L]
                dt = torch.empty(tr.npix, Lp, dtype=ctx.adt, device=dev)[:, :L]
                tr.invalidate(*st.dst)
                bf = K.coupling_bwd(u, du, s, t, IRN, st.clamp, rev, ds, dt, want_bf, fast=ctx.tc)
                st.nets[0].bwd(ctx, tr, saved_s, ds, dsrc)
                st.nets[1].bwd(ctx, tr, saved_t, dt, dsrc)
            else:
                sign = -1.0 if rev else 1.0
                f, saved = stored if stored is not None else st.nets[0].fwd(ctx, tr, st.src, keep=True)
                tr.invalidate(*st.dst)
                K.axpy_slice(u, f, -sign)                       # restore dst
                df = torch.empty(tr.npix, _round_up(L, 8), dtype=ctx.adt, device=dev)[:, :L]
                K.cast_slice(du, df, sign)
                st.nets[0].bwd(ctx, tr, saved, df, dsrc)
                bf = None
            if bf is not None:
                tr.bf[st.dst] = bf
        flush_param_grads(ctx)


def glow_op(channels, s1, s2, clamp):
    """FrEIA GLOWCouplingBlock (call site archs.py:61-64): forward runs s2 on x2 to update x1, then s1 on
    y1 to update x2; rev undoes them in the opposite order."""
    l1 = channels // 2
    h1, h2 = (0, l1), (l1, channels)
    n1, n2 = ConvSubnet(s1), ConvSubnet(s2)
    if n1.cin != l1 or n1.cout != 2 * (channels - l1) or n2.cin != channels - l1 or n2.cout != 2 * l1:
        raise SininnError("GLOW subnet channel counts do not match the split")
    return CouplingOp([HalfStep("glow", [n2], h2, h1, clamp), HalfStep("glow", [n1], h1, h2, clamp)], channels)


def irn_op(channels, split1, F, G, H, clamp):
    """InvBlockExp (archs.py:135-160): y1 = x1 + F(x2); y2 = x2*exp(clamp*(2*sigmoid(H(y1))-1)) + G(y1)."""
    h1, h2 = (0, split1), (split1, channels)
    nF, nG, nH = DenseSubnet(F), DenseSubnet(G), DenseSubnet(H)
    return CouplingOp([HalfStep("irn_add", [nF], h2, h1, clamp), HalfStep("irn_affine", [nH, nG], h1, h2, clamp)],
                      channels)


# ----------------------------------------------------------------------------- the plan
class Plan:
    """Ops in forward order.  Resamples ahead of the first coupling run on NCHW data (few channels);
    from the first coupling on, everything is channels-last.  A trailing permutation is folded into the
    layout change at the API boundary."""

    def __init__(self, ops, in_dims):
        self.ops = ops
        self.in_dims = tuple(in_dims)
        first = next((i for i, o in enumerate(ops) if o.kind == "coupling"), len(ops))
        self.prefix = ops[:first]
        self.body = ops[first:]
        if any(o.kind != "resample" for o in self.prefix):
            first = next(i for i, o in enumerate(ops) if o.kind != "resample")
            self.prefix, self.body = ops[:first], ops[first:]
        # SRF: squeeze, squeeze, first coupling block (archs.py:28-38) -- the two squeezes and the layout change are one kernel
        self.squeeze2 = (FUSE_SQUEEZE2 and bool(self.body) and len(self.prefix) == 2
                         and all(o.kind == "resample" and o.mode == 0 for o in self.prefix) and 16 * self.in_dims[0] <= 128)
        self.tail_perm = self.body[-1] if self.body and self.body[-1].kind == "perm" else None
        self.core = self.body[:-1] if self.tail_perm is not None else self.body
        c, h, w = self.in_dims
        for o in ops:
            if o.kind == "resample":
                c, h, w = c * 4, h // 2, w // 2
        self.out_dims = (c, h, w)
        self._packsets = {}
        # opt-in (train.SingleVideoTrainer): weight-gradient kernels accumulate straight into param.grad and autograd
        # receives None for those parameters -- saves one elementwise add per parameter per backward pass
        self.direct_grad = False
        # opt-in with direct_grad: weight/bias gradient launches run on a side stream next to the data-gradient chain
        self.side_wgrad = False
        self._wstreams = {}
        self._store_choice = {}

    def _wgrad_stream(self):
        cur = torch.cuda.current_stream()
        key = cur.cuda_stream
        if key not in self._wstreams:
            self._wstreams[key] = torch.cuda.Stream(device=cur.device)
        return self._wstreams[key]

    def parameters(self):
        out, seen = [], set()
        for o in self.ops:
            for p in o.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out

    def packs(self, dtype, split=False):
        """The plan's conv weights packed for `dtype`, refreshed (one launch) if any weight changed."""
        key = (dtype, split)
        ps = self._packsets.get(key)
        if ps is None:
            convs = [p for o in self.ops if o.kind == "coupling" for p in o.parameters() if p.dim() == 4]
            ps = self._packsets[key] = PackSet(convs, dtype, split)
        ps.ensure()
        return ps

    def _ctx_packs(self, cfg, t):
        if not (self.body and K.__name__ == "sin_inn_b200.kernels" and t.is_cuda):
            return None, None
        return self.packs(cfg.act_dtype), (self.packs(torch.bfloat16, True) if cfg.split else None)

    def _check_input(self, x, rev):
        if isinstance(x, LatentInput):
            if not rev or not self.body:
                raise SininnError("a LatentInput (lr, z) is the input of the inverse pass of a network with coupling blocks")
        require_cuda(x.lr if isinstance(x, LatentInput) else x, "network input")
        if x.dtype != torch.float32 or x.dim() != 4:
            raise SininnError(f"network input must be a 4-D fp32 NCHW tensor, got {x.dtype} {tuple(x.shape)}")
        want = self.out_dims if rev else self.in_dims
        if x.shape[1] != want[0]:
            raise SininnError(f"expected {want[0]} channels for rev={rev}, got {x.shape[1]}")
        f = 1
        for o in self.ops:
            if o.kind == "resample":
                f *= 2
        if not rev and (x.shape[2] % f or x.shape[3] % f):
            raise SininnError(f"input height/width must be multiples of {f}, got {tuple(x.shape[2:])}")

    def _can_squeeze2(self, t):
        return self.squeeze2 and t.shape[2] % 4 == 0 and t.shape[3] % 4 == 0 and t.data_ptr() % 16 == 0 and t.is_contiguous()

    def _enter(self, x, hint=None):
        """Full-resolution NCHW tensor -> channels-last trunk at the first coupling block's level (+ bf16 hint copy)."""
        if self._can_squeeze2(x):
            return K.squeeze2_to_nhwc(x, hint)
        for op in self.prefix:
            x = op.apply_nchw(x, False)
        return K.nchw_to_nhwc(x, None, hint)

    def _leave(self, U):
        """Inverse of _enter: channels-last trunk -> full-resolution NCHW tensor."""
        if self.squeeze2 and U.is_contiguous():
            return K.nhwc_to_unsqueeze2(U)
        y = K.nhwc_to_nchw(U, None)
        for op in self.prefix[::-1]:
            y = op.apply_nchw(y, True)
        return y

    def _hint(self, op, rev, ctx):
        """bf16 operand range the next coupling will read first (so the producer can emit it)."""
        if op is None or op.kind != "coupling" or ctx.adt != torch.bfloat16:
            return None
        src = op.first_src(rev)
        return src if (src[1] - src[0]) % 8 == 0 and not isinstance(
            (op.steps[-1] if rev else op.steps[0]).nets[0], DenseSubnet) else None

    # ---- value pass ---------------------------------------------------------------------------
    def stash_bytes(self, shape, rev, cfg):
        """Bytes a differentiable value pass keeps in "store" mode for an input of `shape` (per-pixel figures of the
        subnets x the pixel count of the level each coupling block runs at)."""
        npix = shape[0] * shape[2] * shape[3]
        if not rev:
            npix //= 4 ** len(self.prefix)
        total = 0
        for op in (self.core[::-1] if rev else self.core):
            if op.kind == "resample":
                npix = npix * 4 if rev else npix // 4
            elif op.kind == "coupling":
                total += npix * op.stash_bytes_per_pixel(cfg)
        return total

    def wants_store(self, x, rev, cfg):
        """EngineConfig.activations resolved for this call: True = keep the subnet internals of the value pass."""
        if cfg.activations not in ACTIVATION_MODES:
            raise SininnError(f"EngineConfig.activations must be one of {ACTIVATION_MODES}, got {cfg.activations!r}")
        if cfg.activations != "auto" or not self.body:
            return cfg.activations == "store" and bool(self.body)
        key = (tuple(x.shape), bool(rev), cfg.precision, cfg.tensor_core)
        hit = self._store_choice.get(key)
        if hit is None:
            dev = x.device
            if dev.type != "cuda":
                return True
            avail = torch.cuda.get_device_properties(dev).total_memory - torch.cuda.memory_allocated(dev)
            hit = self.stash_bytes(x.shape, rev, cfg) <= STORE_FRACTION * avail
            self._store_choice[key] = hit
        return hit

    def execute(self, x, rev, cfg, stash=None):
        """stash: a list to fill with the subnet internals a later backward(..., stash=) consumes ("store" mode)."""
        self._check_input(x, rev)
        p_plain, p_split = self._ctx_packs(cfg, x)
        ctx = RunCtx(cfg, packs=p_plain, split_packs=p_split, stash=stash)
        if not self.body:
            seq = self.prefix[::-1] if rev else self.prefix
            for op in seq:
                x = op.apply_nchw(x, rev)
            return x
        dev = x.device
        if x.is_cuda and isinstance(x, LatentInput) and x.z is not None and x.z.device != dev:
            raise SininnError("LatentInput: lr and z live on different devices")
        if not rev:
            seq = self.core
            hint = self._hint(seq[0] if seq else None, rev, ctx)
            U, bf = self._enter(x, hint)
        else:
            seq = self.core[::-1]
            cmap = self.tail_perm.gather_map(dev, True) if self.tail_perm is not None else None
            hint = self._hint(seq[0] if seq else None, rev, ctx)
            if isinstance(x, LatentInput):
                U, bf = K.latent_to_nhwc(x.lr, x.z, x.z_dims, cmap, hint, seed=x.seed, offset=x.offset, temp=x.temp,
                                         step_state=x.step_state)
            else:
                U, bf = K.nchw_to_nhwc(x, cmap, hint)
        tr = Trunk(U)
        if bf is not None:
            tr.bf[hint] = bf
        _trace("to_nhwc", tr.U)
        folded = False
        for i, op in enumerate(seq):
            nxt = seq[i + 1] if i + 1 < len(seq) else None
            if folded:                     # this permutation was applied by the coupling block before it
                folded = False
                _trace(op.kind, tr.U)
                continue
            if op.kind == "coupling":
                if nxt is not None and nxt.kind == "perm" and op.folds_perm_value(ctx, tr, rev):
                    after = seq[i + 2] if i + 2 < len(seq) else None
                    op.run(ctx, tr, rev, post_perm=(nxt.gather_map(dev, rev), self._hint(after, rev, ctx)))
                    folded = True
                else:
                    op.run(ctx, tr, rev)
            elif op.kind == "perm":
                hint = self._hint(nxt, rev, ctx)
                U, bf = K.permute_nhwc(tr.U, op.gather_map(dev, rev), hint)
                tr.set(U, None, {hint: bf} if bf is not None else None)
            elif op.kind == "linear":
                tr.set(op.apply(tr.U, rev, split=ctx.split))
            elif op.kind == "actnorm":
                op.apply(tr, rev)
            else:
                tr.set(op.apply_nhwc(tr.U, rev))
            _trace(op.kind, tr.U)
        if not rev:
            cmap = self.tail_perm.gather_map(dev, False) if self.tail_perm is not None else None
            return K.nhwc_to_nchw(tr.U, cmap)
        return self._leave(tr.U)

    # ---- backward from the output ----------------------------------------------------------------
    def backward(self, y, dy, rev, cfg, need_dx=True, stash=None):
        """y: the output execute(x, rev) produced; dy: dL/dy; stash: what execute(..., stash=) kept, consumed here.
        Returns (dL/dx or None, {id(param): grad})."""
        require_cuda(dy, "grad_output")
        p_plain, p_split = self._ctx_packs(cfg, dy)
        ctx = RunCtx(cfg, want_grads=True, direct_grad=self.direct_grad, packs=p_plain, split_packs=p_split, stash=stash)
        dy = dy.contiguous()
        if dy.dtype != torch.float32:
            raise SininnError("grad_output must be fp32")
        if not self.body:
            seq = self.prefix if rev else self.prefix[::-1]
            for op in seq:
                dy = op.apply_nchw(dy, rev, grad=True)
            return dy, ctx.grads
        dev = y.device
        # 1. bring (y, dy) back into the channels-last trunk the value pass ended with
        if not rev:
            cmap = self.tail_perm.gather_map(dev, True) if self.tail_perm is not None else None   # undo perm
            U, _ = K.nchw_to_nhwc(y, cmap, None)
            dU, _ = K.nchw_to_nhwc(dy, cmap, None)
            undo = self.core[::-1]
        else:
            # the value pass ended with inverse resamples on NCHW: re-apply the forward map to get back to the trunk (for
            # the squeeze, the gradient of the inverse map is the forward map too)
            if self._can_squeeze2(y) and self._can_squeeze2(dy):
                U, _ = K.squeeze2_to_nhwc(y, None)
                dU, _ = K.squeeze2_to_nhwc(dy, None)
            else:
                for op in self.prefix:
                    y = op.apply_nchw(y, False)
                    dy = op.apply_nchw(dy, True, grad=True)
                U, _ = K.nchw_to_nhwc(y, None, None)
                dU, _ = K.nchw_to_nhwc(dy, None, None)
            undo = self.core                       # executed order was reversed(core)
        tr = Trunk(U, dU)
        # side stream only on the tensor-core path: its operands are private bf16 copies, while the fp32 path reads views
        # of the live trunk that the next half-step overwrites in place (record_stream does not order that write)
        if self.side_wgrad and self.direct_grad and dy.is_cuda and cfg.tc and K.__name__ == "sin_inn_b200.kernels":
            ctx.wstream = self._wgrad_stream()
        # 2. walk the executed ops backwards
        pre_perm = None
        for i, op in enumerate(undo):
            if op.kind == "coupling":
                op.backward(ctx, tr, rev, pre_perm=pre_perm)
                pre_perm = None
            elif op.kind == "perm":
                m_val = op.gather_map(dev, not rev)            # inverse of the executed value map
                m_grad = op.gather_map(dev, rev, grad=True)
                nxt_op = undo[i + 1] if i + 1 < len(undo) else None
                if (m_val is m_grad and nxt_op is not None and nxt_op.kind == "coupling"
                        and nxt_op.folds_perm_backward(ctx, tr, rev)):
                    pre_perm = m_val                           # undone by the next block's first backward kernel
                    continue
                # the gather also emits the bf16 operand of the subnet the next block's backward evaluates first
                hint = None
                nxt = undo[i + 1] if i + 1 < len(undo) else None
                if nxt is not None and nxt.kind == "coupling" and ctx.adt == torch.bfloat16:
                    st = nxt.first_step_backward(rev)
                    if (st.src[1] - st.src[0]) % 8 == 0 and st.src[0] % 4 == 0 and not isinstance(st.nets[0], DenseSubnet):
                        hint = st.src
                if m_val is m_grad:                            # always: both undo the same gather
                    U, dU, bf = K.permute_nhwc_pair(tr.U, tr.dU, m_val, hint)
                else:
                    U, bf = K.permute_nhwc(tr.U, m_val, hint)
                    dU, _ = K.permute_nhwc(tr.dU, m_grad, None)
                tr.set(U, dU, {hint: bf} if bf is not None else None)
            elif op.kind == "linear":
                # executed y = A x (A = W or W^-1): the input is A^-1 y, its gradient A^T dy
                tr.set(op.apply(tr.U, not rev, split=ctx.split), op.apply(tr.dU, rev, grad=True, split=ctx.split))
            elif op.kind == "actnorm":
                op.backward(ctx, tr, rev)
            else:
                U = op.apply_nhwc(tr.U, not rev)
                dU = op.apply_nhwc(tr.dU, rev, grad=True)
                tr.set(U, dU)
        flush_param_grads(ctx)
        if ctx.wstream is not None:
            torch.cuda.current_stream().wait_stream(ctx.wstream)       # parameter gradients complete with the pass
        if not need_dx:
            return None, ctx.grads
        # 3. gradient back out through the API-side layout change and the NCHW resamples
        if not rev:
            if self.squeeze2 and tr.dU.is_contiguous():
                dx = K.nhwc_to_unsqueeze2(tr.dU)       # gradient of the squeeze = its inverse map
            else:
                dx = K.nhwc_to_nchw(tr.dU, None)
                for op in self.prefix[::-1]:
                    dx = op.apply_nchw(dx, False, grad=True)
        else:
            cmap = self.tail_perm.gather_map(dev, True, grad=True) if self.tail_perm is not None else None
            dx = K.nhwc_to_nchw(tr.dU, cmap)
        return dx, ctx.grads


class _INNFunction(torch.autograd.Function):
    """Differentiable net(x) / net(x, rev=True) that keeps only its output for backward."""

    @staticmethod
    def forward(ctx, x, plan, rev, cfg, latent, *params):
        inp = latent if latent is not None else x.detach()
        ctx.stash = [] if plan.wants_store(inp, rev, cfg) else None
        y = plan.execute(inp, rev, cfg, stash=ctx.stash)
        ctx.plan, ctx.rev, ctx.cfg = plan, rev, cfg
        ctx.params = params
        ctx.need_dx = x.requires_grad and latent is None
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        stash, ctx.stash = ctx.stash, None             # consumed: a second backward through this node is not supported
        with _device_ctx(y):
            dx, grads = ctx.plan.backward(y, dy, ctx.rev, ctx.cfg, need_dx=ctx.need_dx, stash=stash)
        gl = [grads.get(id(p)) if p.requires_grad else None for p in ctx.params]
        return (dx, None, None, None, None, *gl)


def _device_ctx(t):
    return torch.cuda.device(t.device)


def run_network(plan, x, rev, cfg):
    """net(x) / net(x, rev=True).  Everything is launched on x's device and on that device's current stream: the
    kernels take raw pointers and the library uses the current CUDA device (stream, SM count, function attributes),
    so the device is made current for the duration of the call (the reference loads checkpoints onto cuda:{gpu_ids[0]}
    while another device may be current, main.py:127)."""
    latent = x if isinstance(x, LatentInput) else None
    if latent is not None:
        x = latent.lr
    require_cuda(x, "network input")
    params = plan.parameters()
    for p in params:
        if p.device != x.device:
            raise SininnError(f"network parameters live on {p.device} but the input is on {x.device}")
    needs_grad = torch.is_grad_enabled() and ((x.requires_grad and latent is None) or any(p.requires_grad for p in params))
    with _device_ctx(x):
        if not needs_grad:
            return plan.execute(latent if latent is not None else x, rev, cfg)
        return _INNFunction.apply(x, plan, bool(rev), cfg, latent, *params)
