"""Checks run both on the CPU stand-in kernels (host tests) and on the real kernels (GPU tests)."""
import torch

from sin_inn_b200 import archs, engine as E


def fixed1x1_checks(dev):
    """Shared by the host (fake kernels) and GPU tests: Fixed1x1Conv against torch's conv2d, exact round trip,
    log-determinant, gradient, the permutation-matrix special case, and a chain squeeze -> 1x1 -> GLOW."""
    import numpy as np
    import torch.nn.functional as F
    from sin_inn_b200.freia import framework as Ff, modules as Fm
    C, h, w = 12, 6, 10
    g = torch.Generator().manual_seed(4)
    M = torch.linalg.qr(torch.randn(C, C, generator=g, dtype=torch.float64))[0] * 1.3 + 0.05 * torch.randn(C, C, generator=g, dtype=torch.float64)
    x = torch.randn(2, C, h, w, generator=g)
    m = Fm.Fixed1x1Conv([(C, h, w)], M).to(dev)
    xg = x.to(dev).requires_grad_(True)
    y = m([xg])[0]
    ref = F.conv2d(x.double(), M.t().reshape(C, C, 1, 1)).float()
    assert y.is_contiguous() and (y.detach().cpu() - ref).abs().max() < 2e-6 * ref.abs().max()
    back = m([y.detach()], rev=True)[0]
    assert (back.cpu() - x).abs().max() < 1e-5
    want_ld = float(torch.linalg.slogdet(M)[1]) * h * w
    assert torch.allclose(m.jacobian([xg]).cpu(), torch.full((2,), want_ld), rtol=1e-6)
    assert torch.allclose(m.jacobian([xg], rev=True).cpu(), torch.full((2,), -want_ld), rtol=1e-6)
    (y * y).sum().backward()
    xr = x.double().clone().detach().requires_grad_(True)
    (F.conv2d(xr, M.t().reshape(C, C, 1, 1)) ** 2).sum().backward()
    assert (xg.grad.cpu() - xr.grad.float()).abs().max() < 1e-5 * xr.grad.abs().max()
    # a permutation matrix reproduces PermuteRandom (its special case)
    pm = Fm.PermuteRandom([(C, h, w)], seed=2).to(dev)
    P = torch.zeros(C, C, dtype=torch.float64)
    P[pm.perm, torch.arange(C)] = 1.0                      # y[:, o] = x[:, perm[o]]
    mp = Fm.Fixed1x1Conv([(C, h, w)], P).to(dev)
    assert torch.equal(mp([x.to(dev)])[0], pm([x.to(dev)])[0]) and abs(mp._lin.logdet) < 1e-12
    # inside a graph: squeeze -> fixed 1x1 -> GLOW -> permute, the layout the reference sketches at archs.py:40-68
    torch.manual_seed(1)
    nodes = [Ff.InputNode(3, 8, 8, name="in")]
    nodes.append(Ff.Node(nodes[-1], Fm.IRevNetDownsampling, {}, name="sq"))
    Q = torch.linalg.qr(torch.randn(12, 12, generator=g, dtype=torch.float64))[0]
    nodes.append(Ff.Node(nodes[-1], Fm.Fixed1x1Conv, {"M": Q}, name="conv_1x1"))
    nodes.append(Ff.Node(nodes[-1], Fm.GLOWCouplingBlock, {"subnet_constructor": archs.subnet_conv_1x1, "clamp": 1.2}, name="glow"))
    nodes.append(Ff.Node(nodes[-1], Fm.PermuteRandom, {"seed": 0}, name="perm"))
    nodes.append(Ff.OutputNode(nodes[-1], name="out"))
    net = Ff.ReversibleGraphNet(nodes, verbose=False).to(dev)
    img0 = torch.rand(2, 3, 8, 8, generator=g).to(dev)
    d = torch.randn(img0.shape, generator=g).to(dev)
    outs = {}
    # "fp32": the 1x1 convolution on the CUDA-core fp32 GEMM; "fp32tc": on the tensor cores over split operands (north-star
    # item 3: the invertible 1x1 convolution as a tcgen05 GEMM, at fp32 accuracy)
    for precision in ("fp32", "fp32tc"):
        net.engine_config = E.EngineConfig(precision=precision)
        net.zero_grad()
        img = img0.clone().requires_grad_(True)
        out = net(img)
        outs[precision] = out.detach()
        assert (net(out.detach(), rev=True) - img.detach()).abs().max() < 1e-5
        out.square().mean().backward()
        # gradient check against finite differences of the same network
        with torch.no_grad():
            eps = 1e-2
            num = (net(img.detach() + eps * d).square().mean() - net(img.detach() - eps * d).square().mean()) / (2 * eps)
        assert abs(float(num) - float((img.grad * d).sum())) < 2e-3 * max(1.0, abs(float(num)))
    assert (outs["fp32"] - outs["fp32tc"]).abs().max() < 1e-4 * max(1.0, float(outs["fp32"].abs().max()))
