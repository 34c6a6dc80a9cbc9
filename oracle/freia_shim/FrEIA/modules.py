"""ORACLE ONLY. The three FrEIA operators the reference names
(/root/reference/archs.py:28-31,35-38 IRevNetDownsampling; :61-64
GLOWCouplingBlock(subnet_constructor=, clamp=); :65-68 PermuteRandom(seed=)),
restated from the published pre-v0.2 semantics."""
import numpy as np
import torch
import torch.nn as nn


class IRevNetDownsampling(nn.Module):
    """Legacy i-RevNet space-to-depth:
    out[b,(dy*2+dx)*C+c,i,j] = in[b,c,2i+dy,2j+dx]   (NOT pixel_unshuffle order)."""

    def __init__(self, dims_in):
        super().__init__()
        self.block_size = 2

    def forward(self, x, rev=False):
        t = x[0]
        b, c, h, w = t.shape
        if not rev:
            v = t.reshape(b, c, h // 2, 2, w // 2, 2)          # b c i dy j dx
            v = v.permute(0, 3, 5, 1, 2, 4)                    # b dy dx c i j
            return [v.reshape(b, 4 * c, h // 2, w // 2).contiguous()]
        co = c // 4
        v = t.reshape(b, 2, 2, co, h, w)                       # b dy dx c i j
        v = v.permute(0, 3, 4, 1, 5, 2)                        # b c i dy j dx
        return [v.reshape(b, co, 2 * h, 2 * w).contiguous()]

    def jacobian(self, x, rev=False):
        return 0

    def output_dims(self, input_dims):
        c, h, w = input_dims[0]
        return [(c * 4, h // 2, w // 2)]


class GLOWCouplingBlock(nn.Module):
    """Affine coupling, soft clamp e(s)=exp(clamp*0.636*atan(s/clamp)) -- the
    literal 0.636, and s/clamp inside atan (pre-v0.2).  s1 is constructed before
    s2 (fixes RNG order and the state_dict keys ``s1.*``/``s2.*``)."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor=None, clamp=5.0):
        super().__init__()
        channels = dims_in[0][0]
        self.ndims = len(dims_in[0])
        self.split_len1 = channels // 2
        self.split_len2 = channels - channels // 2
        self.clamp = clamp
        self.s1 = subnet_constructor(self.split_len1, self.split_len2 * 2)
        self.s2 = subnet_constructor(self.split_len2, self.split_len1 * 2)

    def log_e(self, s):
        return self.clamp * 0.636 * torch.atan(s / self.clamp)

    def e(self, s):
        return torch.exp(self.log_e(s))

    def forward(self, x, c=[], rev=False):
        x1 = x[0].narrow(1, 0, self.split_len1)
        x2 = x[0].narrow(1, self.split_len1, self.split_len2)
        dims = tuple(range(1, self.ndims + 1))
        if not rev:
            r2 = self.s2(x2)
            s2, t2 = r2[:, :self.split_len1], r2[:, self.split_len1:]
            y1 = self.e(s2) * x1 + t2
            r1 = self.s1(y1)
            s1, t1 = r1[:, :self.split_len2], r1[:, self.split_len2:]
            y2 = self.e(s1) * x2 + t1
            self.last_jac = self.log_e(s1).sum(dims) + self.log_e(s2).sum(dims)
        else:
            r1 = self.s1(x1)
            s1, t1 = r1[:, :self.split_len2], r1[:, self.split_len2:]
            y2 = (x2 - t1) / self.e(s1)
            r2 = self.s2(y2)
            s2, t2 = r2[:, :self.split_len1], r2[:, self.split_len1:]
            y1 = (x1 - t2) / self.e(s2)
            self.last_jac = -self.log_e(s1).sum(dims) - self.log_e(s2).sum(dims)
        return [torch.cat((y1, y2), 1)]

    def jacobian(self, x, c=[], rev=False):
        return self.last_jac

    def output_dims(self, input_dims):
        return input_dims


class PermuteRandom(nn.Module):
    """Fixed channel permutation from numpy's legacy MT19937 stream:
    np.random.seed(seed); perm = np.random.permutation(C); fwd x[:, perm],
    rev x[:, perm_inv] with perm_inv[perm[i]] = i.  Tables are plain attributes
    (not buffers, not in the state_dict)."""

    def __init__(self, dims_in, seed):
        super().__init__()
        self.in_channels = dims_in[0][0]
        np.random.seed(seed)
        perm = np.random.permutation(self.in_channels)
        np.random.seed()
        perm_inv = np.zeros_like(perm)
        for i, p in enumerate(perm):
            perm_inv[p] = i
        self.perm = torch.LongTensor(perm)
        self.perm_inv = torch.LongTensor(perm_inv)

    def forward(self, x, rev=False):
        if not rev:
            return [x[0][:, self.perm]]
        return [x[0][:, self.perm_inv]]

    def jacobian(self, x, rev=False):
        return 0.0

    def output_dims(self, input_dims):
        return input_dims


class ActNorm(nn.Module):
    """Per-channel affine normalisation with data-dependent initialisation (the node the reference leaves commented
    out at /root/reference/archs.py:40-44), restated from the published pre-v0.2 source: scale/bias are [1,C,1,1]
    parameters; the first batch sets scale = log(1/std_c), bias = -mean_c(x * exp(scale)); forward x*exp(scale)+bias;
    jacobian = sum(scale) * H*W per sample."""

    def __init__(self, dims_in, init_data=None):
        super().__init__()
        self.dims_in = dims_in[0]
        param_dims = [1, self.dims_in[0]] + [1 for _ in range(len(self.dims_in) - 1)]
        self.scale = nn.Parameter(torch.zeros(*param_dims))
        self.bias = nn.Parameter(torch.zeros(*param_dims))
        self.init_on_next_batch = True
        if init_data is not None:
            self.initialize_with_data(init_data)

        def on_load_state_dict(*args):
            self.init_on_next_batch = False
        self._register_load_state_dict_pre_hook(on_load_state_dict)

    def initialize_with_data(self, data):
        assert all(data.shape[i + 1] == self.dims_in[i] for i in range(len(self.dims_in)))
        self.scale.data.view(-1)[:] = torch.log(1 / data.transpose(0, 1).contiguous().view(self.dims_in[0], -1).std(dim=-1))
        data = data * self.scale.exp()
        self.bias.data.view(-1)[:] = -data.transpose(0, 1).contiguous().view(self.dims_in[0], -1).mean(dim=-1)
        self.init_on_next_batch = False

    def forward(self, x, rev=False):
        if self.init_on_next_batch:
            self.initialize_with_data(x[0])
        if not rev:
            return [x[0] * self.scale.exp() + self.bias]
        return [(x[0] - self.bias) / self.scale.exp()]

    def jacobian(self, x, rev=False):
        j = self.scale.sum() * np.prod(self.dims_in[1:])
        return (-j if rev else j).repeat(x[0].shape[0])

    def output_dims(self, input_dims):
        return input_dims
