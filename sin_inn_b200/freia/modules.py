"""FrEIA.modules drop-ins (module protocol: ``M(dims_in, **kwargs)``, ``forward([x], rev=False) -> [y]``,
``jacobian(x, rev=False)``, ``output_dims(dims_in)``) backed by libsininn kernels.

Replaces, at their call sites in the reference:
  IRevNetDownsampling   archs.py:28-31, 35-38
  GLOWCouplingBlock     archs.py:61-64   (subnet_constructor=, clamp=)
  PermuteRandom         archs.py:65-68   (seed=)
  Fixed1x1Conv          archs.py:40-50   (commented out there; offered as the north-star extension)
CUDA tensors only -- there is no CPU path.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import engine as E


class _PlanModule(nn.Module):
    """One operator = a one-op plan, so standalone use shares the executor (and its autograd) with whole nets."""

    def _op(self):
        raise NotImplementedError

    def _plan(self):
        if getattr(self, "_plan_cache", None) is None:
            object.__setattr__(self, "_plan_cache", E.Plan([self._op()], self.dims_in))
        return self._plan_cache

    def forward(self, x, c=[], rev=False):
        if isinstance(x, (list, tuple)):
            return [E.run_network(self._plan(), x[0], rev, E.default_config())]
        return E.run_network(self._plan(), x, rev, E.default_config())

    def output_dims(self, input_dims):
        return input_dims


class IRevNetDownsampling(_PlanModule):
    """out[b,(dy*2+dx)*C+c,i,j] = in[b,c,2i+dy,2j+dx] (legacy i-RevNet order)."""

    def __init__(self, dims_in):
        super().__init__()
        self.dims_in = tuple(dims_in[0])
        self.block_size = 2

    def _op(self):
        return E.ResampleOp(0)

    def jacobian(self, x, rev=False):
        return 0

    def output_dims(self, input_dims):
        c, h, w = input_dims[0]
        if h % 2 or w % 2:
            raise E.SininnError(f"IRevNetDownsampling needs even height/width, got {h}x{w}")
        return [(c * 4, h // 2, w // 2)]


class GLOWCouplingBlock(_PlanModule):
    """y1 = e(s2(x2))*x1 + t2(x2); y2 = e(s1(y1))*x2 + t1(y1); e(s) = exp(clamp*0.636*atan(s/clamp))."""

    def __init__(self, dims_in, dims_c=[], subnet_constructor=None, clamp=5.0):
        super().__init__()
        if dims_c:
            raise E.SininnError("conditional coupling blocks are not supported (the reference never uses them)")
        self.dims_in = tuple(dims_in[0])
        channels = self.dims_in[0]
        self.ndims = len(self.dims_in)
        self.split_len1 = channels // 2
        self.split_len2 = channels - channels // 2
        self.clamp = clamp
        # construction order s1, s2 matters: it fixes RNG consumption and the state_dict keys
        self.s1 = subnet_constructor(self.split_len1, self.split_len2 * 2)
        self.s2 = subnet_constructor(self.split_len2, self.split_len1 * 2)

    def _op(self):
        return E.glow_op(self.dims_in[0], self.s1, self.s2, self.clamp)

    def jacobian(self, x, c=[], rev=False):
        """log|det J| per sample of the block evaluated at x in direction `rev` (FrEIA returns the `last_jac` its
        forward stored: sum over channels and pixels of the clamped log-scales of both halves, negated for rev).
        The reference never reads it (loss.py:38-39 uses mean(z^2)); it is computed here on demand by re-running
        the block's value pass with one reduction launch per half."""
        x0 = x[0] if isinstance(x, (list, tuple)) else x
        E.require_cuda(x0, "jacobian input")
        cfg = E.default_config()
        with torch.no_grad(), E._device_ctx(x0):
            U, _ = E.K.nchw_to_nhwc(x0.detach(), None, None)
            tr = E.Trunk(U)
            logdet = torch.zeros(x0.shape[0], dtype=torch.float32, device=x0.device)
            op = self._plan().core[0]
            op.run(E.RunCtx(cfg, packs=None), tr, rev, logdet=logdet)
        return logdet


class PermuteRandom(_PlanModule):
    """np.random.seed(seed); perm = np.random.permutation(C); fwd x[:, perm], rev x[:, perm_inv]."""

    def __init__(self, dims_in, seed):
        super().__init__()
        self.dims_in = tuple(dims_in[0])
        self.in_channels = self.dims_in[0]
        np.random.seed(seed)
        perm = np.random.permutation(self.in_channels)
        np.random.seed()
        inv = np.zeros_like(perm)
        inv[perm] = np.arange(self.in_channels)
        # plain attributes as upstream: not buffers, not in the state_dict
        self.perm = torch.LongTensor(perm)
        self.perm_inv = torch.LongTensor(inv)

    def _op(self):
        return E.PermOp(self.perm)

    def jacobian(self, x, rev=False):
        return 0.0


class ActNorm(_PlanModule):
    """Per-channel affine normalisation with data-dependent initialisation: forward x * exp(scale) + bias, reverse
    (x - bias) / exp(scale); log|det J| = H*W * sum(scale).  Offered (commented out) at archs.py:40-44."""

    def __init__(self, dims_in, init_data=None):
        super().__init__()
        self.dims_in = tuple(dims_in[0])
        param_dims = [1, self.dims_in[0]] + [1] * (len(self.dims_in) - 1)
        self.scale = nn.Parameter(torch.zeros(*param_dims))
        self.bias = nn.Parameter(torch.zeros(*param_dims))
        self.init_on_next_batch = True
        if init_data is not None:
            self.initialize_with_data(init_data)

        def on_load_state_dict(*args):
            self.init_on_next_batch = False
        self._register_load_state_dict_pre_hook(on_load_state_dict)

    def initialize_with_data(self, data):
        with torch.no_grad():
            flat = data.transpose(0, 1).contiguous().view(self.dims_in[0], -1)
            self.scale.data.view(-1)[:] = torch.log(1 / flat.std(dim=-1))
            scaled = data * self.scale.exp()
            self.bias.data.view(-1)[:] = -scaled.transpose(0, 1).contiguous().view(self.dims_in[0], -1).mean(dim=-1)
        self.init_on_next_batch = False

    def _op(self):
        return E.ActNormOp(self)

    def jacobian(self, x, rev=False):
        x0 = x[0] if isinstance(x, (list, tuple)) else x
        j = self.scale.detach().sum() * (x0.shape[2] * x0.shape[3])
        return (-j if rev else j).repeat(x0.shape[0])


class Fixed1x1Conv(_PlanModule):
    """Fixed invertible 1x1 convolution: forward conv2d(x, M^T as a [C,C,1,1] kernel), i.e. y[:, o] = sum_i M[i, o]
    x[:, i]; reverse with M^-1; log|det J| = +-(H*W) * log|det M|.  The reference leaves this node commented out
    (archs.py:40-50, "How do we compute M"); M is supplied by the caller, e.g. a random rotation."""

    def __init__(self, dims_in, M):
        super().__init__()
        self.dims_in = tuple(dims_in[0])
        M = torch.as_tensor(M)
        if tuple(M.shape) != (self.dims_in[0], self.dims_in[0]):
            raise E.SininnError(f"Fixed1x1Conv: M must be {self.dims_in[0]}x{self.dims_in[0]}, got {tuple(M.shape)}")
        self._lin = E.LinearOp(M)
        # frozen parameters with the upstream names, so checkpoints carry the matrix
        self.M = nn.Parameter(M.t().to(torch.float32).contiguous().view(*M.shape, 1, 1), requires_grad=False)
        self.M_inv = nn.Parameter(self._lin.mats[(True, False)].to(torch.float32).contiguous().view(*M.shape, 1, 1), requires_grad=False)
        self.logDetM = nn.Parameter(torch.tensor(self._lin.logdet, dtype=torch.float32), requires_grad=False)

    def _op(self):
        return self._lin

    def jacobian(self, x, rev=False):
        x0 = x[0] if isinstance(x, (list, tuple)) else x
        n_pixels = x0.shape[2] * x0.shape[3]
        j = self._lin.logdet * n_pixels
        return torch.full((x0.shape[0],), -j if rev else j, dtype=torch.float32, device=x0.device)
