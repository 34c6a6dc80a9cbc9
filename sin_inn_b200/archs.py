"""Drop-in for the reference's ``archs.py``: same public names, constructor signatures, parameter
names / shapes / initialisation (so reference checkpoints load and ``lit_wrapper.py`` /
``main.py`` work unchanged), computed by the libsininn sm_100a kernels.

  UncondSRFlow(c, h, w, opt)   reference archs.py:19-71    (returns the graph net from __new__)
  InvRescaleNet(c, h, w, opt)  reference archs.py:201-233
  HaarDownsampling / InvBlockExp / DenseBlock               archs.py:162-199 / 135-160 / 74-133
  subnet_fc / subnet_conv / subnet_conv_1x1                 archs.py:7-17

``opt`` fields read: scale, num_coupling, lr_dims (as the reference) plus the optional extensions
``hidden`` (subnet width, reference hard-codes 256), ``precision`` ("bf16" | "fp32" | "fp32tc") and ``activations``
("auto" | "store" | "recompute": what a differentiable pass keeps of the coupling subnets, engine.EngineConfig).
Inputs must be CUDA tensors: there is no CPU implementation in this package.
"""
import numpy as np
import torch
import torch.nn as nn

from . import engine as E
from .freia import framework as Ff
from .freia import modules as Fm

_HIDDEN = 256


def subnet_fc(c_in, c_out):
    return nn.Sequential(nn.Linear(c_in, 512), nn.ReLU(), nn.Linear(512, c_out))


def subnet_conv(c_in, c_out, hidden=_HIDDEN):
    return nn.Sequential(nn.Conv2d(c_in, hidden, 3, padding=1), nn.ReLU(), nn.Conv2d(hidden, c_out, 3, padding=1))


def subnet_conv_1x1(c_in, c_out, hidden=_HIDDEN):
    return nn.Sequential(nn.Conv2d(c_in, hidden, 1), nn.ReLU(), nn.Conv2d(hidden, c_out, 1))


def _levels(scale):
    return (scale - 1).bit_length()


def _config_from(opt):
    prec = getattr(opt, "precision", None)
    if prec is None:
        return None
    return E.EngineConfig(precision=prec, tensor_core=getattr(opt, "tensor_core", True),
                          activations=getattr(opt, "activations", E.default_config().activations))


class UncondSRFlow:
    """Unconditional SRFlow-style INN.  Node order (= state_dict key indices): input, squeeze_init, then per
    level: squeeze_<l>, num_coupling x (glow_<l>_<k>, permute_<l>_<k>); output.  Even k uses the 3x3 subnet,
    odd k the 1x1 subnet; clamp 1.2; permutation seed = k (re-used on every level, as in the reference)."""

    def __new__(cls, c, h, w, opt):
        hidden = getattr(opt, "hidden", _HIDDEN)
        ctors = (lambda ci, co: subnet_conv(ci, co, hidden), lambda ci, co: subnet_conv_1x1(ci, co, hidden))
        chain = [Ff.InputNode(c, h, w, name="input")]

        def push(module_type, args, name):
            chain.append(Ff.Node(chain[-1], module_type, args, name=name))

        push(Fm.IRevNetDownsampling, {}, "squeeze_init")
        for level in range(_levels(opt.scale)):
            push(Fm.IRevNetDownsampling, {}, f"squeeze_{level}")
            for k in range(opt.num_coupling):
                push(Fm.GLOWCouplingBlock, {"subnet_constructor": ctors[k % 2], "clamp": 1.2}, f"glow_{level}_{k}")
                push(Fm.PermuteRandom, {"seed": k}, f"permute_{level}_{k}")
        chain.append(Ff.OutputNode(chain[-1], name="output"))
        net = Ff.ReversibleGraphNet(chain, verbose=False)
        net.engine_config = _config_from(opt)
        return net


class _SubnetFunction(torch.autograd.Function):
    """Standalone differentiable DenseBlock (NCHW in/out) on the engine's subnet kernels."""

    @staticmethod
    def forward(ctx, x, block, cfg, *params):
        K = E.K
        rc = E.RunCtx(cfg)
        U, _ = K.nchw_to_nhwc(x.detach(), None, None)
        tr = E.Trunk(U)
        net = E.DenseSubnet(block)
        out, _ = net.fwd(rc, tr, (0, net.cin))
        B, h, w, _c = U.shape
        ctx.block, ctx.cfg, ctx.params = block, cfg, params
        ctx.save_for_backward(x)
        return K.nhwc_to_nchw(out.view(B, h, w, net.cout), None)

    @staticmethod
    def backward(ctx, dy):
        K = E.K
        (x,) = ctx.saved_tensors
        rc = E.RunCtx(ctx.cfg, want_grads=True)
        U, _ = K.nchw_to_nhwc(x, None, None)
        tr = E.Trunk(U, torch.zeros_like(U))
        net = E.DenseSubnet(ctx.block)
        _, saved = net.fwd(rc, tr, (0, net.cin))
        dU, _ = K.nchw_to_nhwc(dy.contiguous(), None, None)
        dout = torch.empty(tr.npix, net.cout, dtype=rc.adt, device=x.device)
        K.cast_slice(dU.view(tr.npix, net.cout), dout)
        net.bwd(rc, tr, saved, dout, tr.dmat())
        gl = [rc.grads.get(id(p)) if p.requires_grad else None for p in ctx.params]
        return (K.nhwc_to_nchw(tr.dU, None), None, None, *gl)


class DenseBlock(nn.Module):
    """Five 3x3 convs with dense concatenation and LeakyReLU(0.2); conv1-4 xavier-normal * 0.1,
    conv5 zero-initialised (so a fresh InvBlockExp is the identity), biases zero."""

    def __init__(self, channel_in, channel_out, init="xavier", gc=32, bias=True):
        super().__init__()
        widths = [channel_in + i * gc for i in range(5)]
        self.conv1 = nn.Conv2d(widths[0], gc, 3, 1, 1, bias=bias)
        self.conv2 = nn.Conv2d(widths[1], gc, 3, 1, 1, bias=bias)
        self.conv3 = nn.Conv2d(widths[2], gc, 3, 1, 1, bias=bias)
        self.conv4 = nn.Conv2d(widths[3], gc, 3, 1, 1, bias=bias)
        self.conv5 = nn.Conv2d(widths[4], channel_out, 3, 1, 1, bias=bias)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=True)
        if init == "xavier":
            for conv in (self.conv1, self.conv2, self.conv3, self.conv4):
                nn.init.xavier_normal_(conv.weight)
                conv.weight.data *= 0.1
                if conv.bias is not None:
                    conv.bias.data.zero_()
        nn.init.kaiming_normal_(self.conv5.weight, a=0, mode="fan_in")   # consumes RNG exactly like the reference
        self.conv5.weight.data *= 0
        if self.conv5.bias is not None:
            self.conv5.bias.data.zero_()

    def forward(self, x):
        E.require_cuda(x, "DenseBlock input")
        params = [p for p in self.parameters()]
        return _SubnetFunction.apply(x, self, E.default_config(), *params)


class InvBlockExp(Fm._PlanModule):
    def __init__(self, channel_num, channel_split_num, clamp=1.0):
        super().__init__()
        self.split_len1 = channel_split_num
        self.split_len2 = channel_num - channel_split_num
        self.clamp = clamp
        self.dims_in = (channel_num, 0, 0)
        self.F = DenseBlock(self.split_len2, self.split_len1)
        self.G = DenseBlock(self.split_len1, self.split_len2)
        self.H = DenseBlock(self.split_len1, self.split_len2)

    def _op(self):
        return E.irn_op(self.split_len1 + self.split_len2, self.split_len1, self.F, self.G, self.H, self.clamp)

    def forward(self, x, rev=False):
        return E.run_network(self._plan(), x, rev, E.default_config())


class HaarDownsampling(Fm._PlanModule):
    def __init__(self, channel_in):
        super().__init__()
        self.channel_in = channel_in
        self.dims_in = (channel_in, 0, 0)
        k = torch.ones(4, 1, 2, 2)
        k[1, 0, :, 1] = -1          # horizontal detail
        k[2, 0, 1, :] = -1          # vertical detail
        k[3, 0, 1, 0] = -1          # diagonal detail
        k[3, 0, 0, 1] = -1
        # frozen Parameter kept for state_dict compatibility with reference checkpoints; the kernel
        # hard-codes these +-1 patterns
        self.haar_weights = nn.Parameter(torch.cat([k] * channel_in, 0), requires_grad=False)

    def _op(self):
        return E.ResampleOp(1)

    def forward(self, x, rev=False):
        # bookkeeping attributes of the reference (archs.py:184-185, 193-194); nothing reads last_jac
        self.elements = x.shape[1] * x.shape[2] * x.shape[3]
        self.last_jac = self.elements / 4 * np.log(16.0 if rev else 1 / 16.0)
        return E.run_network(self._plan(), x, rev, E.default_config())


class InvRescaleNet(nn.Module):
    """Haar(c), then per level: Haar(C) and num_coupling x InvBlockExp(C, min(lr_dims, C // 2))."""

    def __init__(self, c, h, w, opt):
        super().__init__()
        blocks = [HaarDownsampling(c)]
        channels = 4 * c
        for _ in range(_levels(opt.scale)):
            blocks.append(HaarDownsampling(channels))
            channels *= 4
            blocks.extend(InvBlockExp(channels, min(opt.lr_dims, channels // 2)) for _ in range(opt.num_coupling))
        self.operations = nn.ModuleList(blocks)
        self.in_dims = (c, h, w)
        self.engine_config = _config_from(opt)
        self._net_plan = None

    def plan(self):
        if self._net_plan is None:
            self._net_plan = E.Plan([Ff.op_from_module(m) for m in self.operations], self.in_dims)
        return self._net_plan

    def forward(self, x, rev=False):
        return E.run_network(self.plan(), x, rev, self.engine_config or E.default_config())
