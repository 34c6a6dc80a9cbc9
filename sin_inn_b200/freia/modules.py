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
        raise E.SininnError("log-Jacobian is not tracked: the reference never reads it "
                            "(last_jac is written and never consumed; loss.py:38-39 uses mean(z^2))")


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
