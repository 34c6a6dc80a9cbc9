"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU (plain PyTorch, fp32 or fp64) restatement of the reference's INN hot path
so that it can run on a box where /root/reference does not exist.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.

What is restated, and from where (all paths relative to /root/reference):
  * SRF graph            archs.py:19-71   (build_srf; operators come from the
                                           FrEIA shim in oracle/freia_shim, see
                                           its header: PARITY UNPINNED at the
                                           FrEIA boundary)
  * subnet constructors  archs.py:11-17   (subnet_conv, subnet_conv_1x1)
  * Haar down/up-sample  archs.py:162-199 (HaarDownsampling: explicit butterfly
                                           instead of grouped conv2d)
  * DenseBlock           archs.py:74-133
  * InvBlockExp          archs.py:135-160
  * InvRescaleNet        archs.py:201-233 (build_irn)
  * train step           lit_wrapper.py:36-56,76 with main.py:52-56 defaults
                                           (train_step; loss.mmd excluded: its
                                           lambda is 0 and loss.py:27-29
                                           hard-codes .to('cuda'))
  * losses               loss.py:3-5,38-39

Pinned by tests/golden/*.npz, which oracle/make_golden.py generates by running
the UNMODIFIED /root/reference/archs.py (through the shim) on seeded inputs.
Module/attribute names equal the reference's so a reference state_dict loads.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "freia_shim")


def _freia():
    """Load the oracle's FrEIA shim by file path under oracle-private module names
    (so it can never shadow, or be shadowed by, a product-side ``FrEIA`` alias)."""
    import importlib.util
    mods = []
    for name in ("framework", "modules"):
        key = "_oracle_freia_" + name
        if key not in sys.modules:
            spec = importlib.util.spec_from_file_location(key, os.path.join(_SHIM, "FrEIA", name + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[key] = m
            spec.loader.exec_module(m)
        mods.append(sys.modules[key])
    return mods[0], mods[1]


# ----------------------------------------------------------------------------- subnets
def subnet_conv(c_in, c_out, hidden=256):
    """archs.py:11-13 -- 3x3 conv, ReLU, 3x3 conv (width 256 hard-coded there)."""
    return nn.Sequential(nn.Conv2d(c_in, hidden, 3, padding=1), nn.ReLU(),
                         nn.Conv2d(hidden, c_out, 3, padding=1))


def subnet_conv_1x1(c_in, c_out, hidden=256):
    """archs.py:15-17 -- 1x1 conv, ReLU, 1x1 conv."""
    return nn.Sequential(nn.Conv2d(c_in, hidden, 1), nn.ReLU(),
                         nn.Conv2d(hidden, c_out, 1))


def n_levels(scale):
    """archs.py:33 / :209 -- number of per-level squeezes after the initial one."""
    return (scale - 1).bit_length()


# ----------------------------------------------------------------------------- SRF
def build_srf(c, h, w, opt):
    """archs.py:24-71: squeeze_init, then per level: squeeze, num_coupling x
    [GLOW(3x3 subnet for even k, 1x1 for odd k, clamp 1.2), PermuteRandom(seed=k)]."""
    Ff, Fm = _freia()
    hidden = getattr(opt, "hidden", 256)
    nodes = [Ff.InputNode(c, h, w, name="input")]
    nodes.append(Ff.Node(nodes[-1], Fm.IRevNetDownsampling, {}, name="squeeze_init"))
    for ss in range(n_levels(opt.scale)):
        nodes.append(Ff.Node(nodes[-1], Fm.IRevNetDownsampling, {}, name=f"squeeze_{ss}"))
        for kk in range(opt.num_coupling):
            if kk % 2 == 0:
                ctor = (lambda ci, co: subnet_conv(ci, co, hidden))
            else:
                ctor = (lambda ci, co: subnet_conv_1x1(ci, co, hidden))
            nodes.append(Ff.Node(nodes[-1], Fm.GLOWCouplingBlock,
                                 {"subnet_constructor": ctor, "clamp": 1.2}, name=f"glow_{ss}_{kk}"))
            nodes.append(Ff.Node(nodes[-1], Fm.PermuteRandom, {"seed": kk}, name=f"permute_{ss}_{kk}"))
    nodes.append(Ff.OutputNode(nodes[-1], name="output"))
    return Ff.ReversibleGraphNet(nodes, verbose=False)


# ----------------------------------------------------------------------------- IRN
class HaarDownsampling(nn.Module):
    """archs.py:162-199.  Forward: the four 2x2 +-1 Haar patterns (archs.py:167-176)
    applied depthwise with stride 2 and divided by 4, output channel order
    band-major out[k*C+c] (archs.py:188-190).  Reverse: exact inverse (transpose
    of the same patterns, no /4; archs.py:195-199)."""

    def __init__(self, channel_in):
        super().__init__()
        self.channel_in = channel_in
        w = torch.ones(4, 1, 2, 2)
        w[1, 0, 0, 1] = -1; w[1, 0, 1, 1] = -1
        w[2, 0, 1, 0] = -1; w[2, 0, 1, 1] = -1
        w[3, 0, 1, 0] = -1; w[3, 0, 0, 1] = -1
        # kept only so the state_dict has the reference's ``haar_weights`` key
        self.haar_weights = nn.Parameter(torch.cat([w] * channel_in, 0), requires_grad=False)

    def forward(self, x, rev=False):
        if not rev:
            a = x[:, :, 0::2, 0::2]; b = x[:, :, 0::2, 1::2]
            c = x[:, :, 1::2, 0::2]; d = x[:, :, 1::2, 1::2]
            o0 = (a + b + c + d) / 4.0
            o1 = (a - b + c - d) / 4.0
            o2 = (a + b - c - d) / 4.0
            o3 = (a - b - c + d) / 4.0
            return torch.cat((o0, o1, o2, o3), 1)
        C = self.channel_in
        o0, o1, o2, o3 = x[:, :C], x[:, C:2 * C], x[:, 2 * C:3 * C], x[:, 3 * C:]
        B, _, h, w = x.shape
        out = x.new_empty(B, C, 2 * h, 2 * w)
        out[:, :, 0::2, 0::2] = o0 + o1 + o2 + o3
        out[:, :, 0::2, 1::2] = o0 - o1 + o2 - o3
        out[:, :, 1::2, 0::2] = o0 + o1 - o2 - o3
        out[:, :, 1::2, 1::2] = o0 - o1 - o2 + o3
        return out


class DenseBlock(nn.Module):
    """archs.py:74-95: five 3x3 convs with dense concatenation, LeakyReLU(0.2) on
    the first four; conv1-4 xavier-normal x0.1, conv5 zero (archs.py:84-86)."""

    def __init__(self, channel_in, channel_out, gc=32):
        super().__init__()
        cin = channel_in
        self.conv1 = nn.Conv2d(cin, gc, 3, 1, 1)
        self.conv2 = nn.Conv2d(cin + gc, gc, 3, 1, 1)
        self.conv3 = nn.Conv2d(cin + 2 * gc, gc, 3, 1, 1)
        self.conv4 = nn.Conv2d(cin + 3 * gc, gc, 3, 1, 1)
        self.conv5 = nn.Conv2d(cin + 4 * gc, channel_out, 3, 1, 1)
        for m in (self.conv1, self.conv2, self.conv3, self.conv4):
            nn.init.xavier_normal_(m.weight)
            m.weight.data *= 0.1
            m.bias.data.zero_()
        nn.init.kaiming_normal_(self.conv5.weight, a=0, mode="fan_in")
        self.conv5.weight.data *= 0
        self.conv5.bias.data.zero_()

    def forward(self, x):
        feats = [x]
        for conv in (self.conv1, self.conv2, self.conv3, self.conv4):
            feats.append(F.leaky_relu(conv(torch.cat(feats, 1)), 0.2))
        return self.conv5(torch.cat(feats, 1))


class InvBlockExp(nn.Module):
    """archs.py:135-160: y1 = x1 + F(x2); s = clamp*(2*sigmoid(H(y1))-1);
    y2 = x2*exp(s) + G(y1); reverse solves the same equations backwards."""

    def __init__(self, channel_num, channel_split_num, clamp=1.0):
        super().__init__()
        self.split_len1 = channel_split_num
        self.split_len2 = channel_num - channel_split_num
        self.clamp = clamp
        self.F = DenseBlock(self.split_len2, self.split_len1)
        self.G = DenseBlock(self.split_len1, self.split_len2)
        self.H = DenseBlock(self.split_len1, self.split_len2)

    def forward(self, x, rev=False):
        x1 = x[:, :self.split_len1]
        x2 = x[:, self.split_len1:]
        if not rev:
            y1 = x1 + self.F(x2)
            s = self.clamp * (torch.sigmoid(self.H(y1)) * 2 - 1)
            y2 = x2 * torch.exp(s) + self.G(y1)
        else:
            s = self.clamp * (torch.sigmoid(self.H(x1)) * 2 - 1)
            y2 = (x2 - self.G(x1)) / torch.exp(s)
            y1 = x1 - self.F(y2)
        return torch.cat((y1, y2), 1)


class InvRescaleNet(nn.Module):
    """archs.py:201-233: Haar(c), then per level Haar(C) followed by num_coupling
    InvBlockExp(C, min(lr_dims, C//2)); rev walks the list backwards."""

    def __init__(self, c, h, w, opt):
        super().__init__()
        ops = [HaarDownsampling(c)]
        cur = c * 4
        for _ in range(n_levels(opt.scale)):
            ops.append(HaarDownsampling(cur))
            cur *= 4
            for _ in range(opt.num_coupling):
                ops.append(InvBlockExp(cur, min(opt.lr_dims, cur // 2)))
        self.operations = nn.ModuleList(ops)

    def forward(self, x, rev=False):
        seq = self.operations if not rev else reversed(self.operations)
        for op in seq:
            x = op(x, rev)
        return x


def build_irn(c, h, w, opt):
    return InvRescaleNet(c, h, w, opt)


def randomize_irn_conv5(net, seed=1, std=0.02):
    """conv5 is zero-initialised (archs.py:86), which makes every InvBlockExp the
    identity; parity tests and the benchmark draw conv5 ~ N(0, std) so the blocks
    do real work."""
    g = torch.Generator().manual_seed(seed)
    for name, p in net.named_parameters():
        if ".conv5.weight" in name:
            p.data.copy_(torch.randn(p.shape, generator=g) * std)
        elif ".conv5.bias" in name:
            p.data.copy_(torch.randn(p.shape, generator=g) * std)


def build(arch, c, h, w, opt):
    return {"SRF": build_srf, "IRN": build_irn}[arch](c, h, w, opt)


def make_opt(scale=4, num_coupling=4, lr_window=10, **kw):
    """main.py:74-75: lr_dims=(2*lr_window+1)*4, z_dims=scale^2*3*4-lr_dims."""
    lr_dims = (2 * lr_window + 1) * 4
    z_dims = scale * scale * 3 * 4 - lr_dims
    d = dict(scale=scale, num_coupling=num_coupling, lr_window=lr_window, lr_dims=lr_dims,
             z_dims=z_dims, lambda_fwd_rec=1.0, lambda_fwd_mmd=0.0, lambda_latent_nll=0.0,
             lambda_bwd_rec=1.0, lambda_bwd_mmd=0.0, lambda_bwd_tcr=0.0, learning_rate=1e-4,
             adam_betas=(0.9, 0.99), weight_decay=1e-5, temp=0.8, architecture="SRF")
    d.update(kw)
    return types.SimpleNamespace(**d)


# ----------------------------------------------------------------------------- losses / step
def reconstruction(x, y):
    """loss.py:3-5."""
    return torch.mean((x - y) ** 2)


def latent_nll(z):
    """loss.py:38-39."""
    return torch.mean(z ** 2)


def mmd(x, y, rev=False):
    """loss.py:9-36 restated device-agnostically (the reference hard-codes .to('cuda') at loss.py:27-29): multi-kernel
    inverse-multiquadric MMD between the flattened batches."""
    kernels = [(0.2, 0.1), (0.2, 0.5), (0.2, 2)] if rev else [(0.2, 2), (1.5, 2), (3.0, 2)]
    b = x.shape[0]
    xf, yf = x.reshape(b, -1), y.reshape(b, -1)
    xx, yy, xy = xf @ xf.t(), yf @ yf.t(), xf @ yf.t()
    rx = xx.diag().unsqueeze(0).expand_as(xx)
    ry = yy.diag().unsqueeze(0).expand_as(yy)
    dxx = torch.clamp(rx.t() + rx - 2.0 * xx, 0, float("inf"))
    dyy = torch.clamp(ry.t() + ry - 2.0 * yy, 0, float("inf"))
    dxy = torch.clamp(rx.t() + ry - 2.0 * xy, 0, float("inf"))
    XX, YY, XY = torch.zeros_like(xx), torch.zeros_like(xx), torch.zeros_like(xx)
    for C, a in kernels:
        XX = XX + C ** a * ((C + dxx) / a) ** -a
        YY = YY + C ** a * ((C + dyy) / a) ** -a
        XY = XY + C ** a * ((C + dxy) / a) ** -a
    return torch.mean(XX + YY - 2.0 * XY)


def train_step(inn, optimizer, hr, lr, z, opt):
    """lit_wrapper.py:36-56,76 without Lightning and without loss.mmd (lambda 0 and
    CUDA-hard-coded, loss.py:27-29): zero_grad; forward pass + L2(+nll) loss +
    backward; reverse pass + L2 loss + backward; one optimizer step."""
    optimizer.zero_grad()
    lr_z = torch.cat((lr, z), dim=1)
    lr_z_hat = inn(hr)
    fwd_loss = opt.lambda_fwd_rec * reconstruction(lr_z_hat[:, :opt.lr_dims], lr)
    fwd_loss = fwd_loss + opt.lambda_latent_nll * latent_nll(lr_z_hat[:, opt.lr_dims:])
    fwd_loss.backward()
    hr_hat = inn(lr_z, rev=True)
    bwd_loss = opt.lambda_bwd_rec * reconstruction(hr_hat, hr)
    bwd_loss.backward()
    optimizer.step()
    return float(fwd_loss.detach()), float(bwd_loss.detach())


def make_optimizer(inn, opt):
    """lit_wrapper.py:131-138."""
    return torch.optim.Adam(inn.parameters(), lr=opt.learning_rate,
                            betas=tuple(opt.adam_betas), weight_decay=opt.weight_decay)


def synthetic_batch(opt, batch, height, width, seed=0, dtype=torch.float32):
    """HR, LR ~ U[0,1), z ~ N(0,1); LR grid is HR/(2*scale)."""
    g = torch.Generator().manual_seed(seed)
    f = 2 * opt.scale
    hr = torch.rand(batch, 3, height, width, generator=g, dtype=dtype)
    lr = torch.rand(batch, opt.lr_dims, height // f, width // f, generator=g, dtype=dtype)
    z = torch.randn(batch, opt.z_dims, height // f, width // f, generator=g, dtype=dtype)
    return hr, lr, z
