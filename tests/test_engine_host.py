"""Host-side logic (plan compilation, half-step bookkeeping, permutation folding, backward-by-inverse)
checked on CPU against the oracle with the torch stand-in kernels.  The real kernels are checked by
the `-m gpu` tests."""
import types

import pytest
import torch

from oracle import ref_torch as R
from sin_inn_b200 import archs, engine as E


def _pair(arch, scale, nc, lr_window, H, W, seed=5, precision="fp32"):
    opt = R.make_opt(scale=scale, num_coupling=nc, lr_window=lr_window, architecture=arch, precision=precision)
    torch.manual_seed(seed)
    ora = R.build(arch, 3, H, W, opt)
    torch.manual_seed(seed)
    net = {"SRF": archs.UncondSRFlow, "IRN": archs.InvRescaleNet}[arch](3, H, W, opt)
    if arch == "IRN":
        R.randomize_irn_conv5(ora, 1)
        R.randomize_irn_conv5(net, 1)
    return opt, ora, net


@pytest.mark.parametrize("arch,scale,nc,lrw", [("SRF", 2, 2, 1), ("SRF", 4, 2, 10), ("IRN", 2, 1, 1), ("IRN", 4, 1, 10)])
def test_same_seed_same_init_and_keys(arch, scale, nc, lrw):
    _, ora, net = _pair(arch, scale, nc, lrw, 16, 16)
    sa, sb = ora.state_dict(), net.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    net.load_state_dict(sa)


@pytest.mark.parametrize("arch,scale,nc,lrw,H,W,precision", [("SRF", 2, 2, 1, 16, 24, "fp32"), ("SRF", 4, 2, 10, 16, 32, "fp32"),
                                                             ("IRN", 2, 1, 1, 16, 16, "fp32"), ("IRN", 4, 1, 10, 16, 24, "fp32"),
                                                             # split-operand plumbing of the fp32 tensor-core path (six-block
                                                             # products, term-list weight gradients) on the torch stand-ins
                                                             ("SRF", 2, 2, 1, 16, 24, "fp32tc"), ("SRF", 4, 2, 10, 16, 16, "fp32tc")])
def test_forward_inverse_backward_match_oracle(fake_kernels, arch, scale, nc, lrw, H, W, precision):
    opt, ora, net = _pair(arch, scale, nc, lrw, H, W, precision=precision)
    hr, lr, z = R.synthetic_batch(opt, 2, H, W, seed=3)
    lrz = torch.cat((lr, z), 1)
    res = {}
    for tag, m in (("ora", ora), ("net", net)):
        for p in m.parameters():
            p.grad = None
        x = hr.clone().requires_grad_(True)
        y = m(x)
        (R.reconstruction(y[:, :opt.lr_dims], lr) + 0.3 * R.latent_nll(y[:, opt.lr_dims:])).backward()
        u = lrz.clone().requires_grad_(True)
        xr = m(u, rev=True)
        R.reconstruction(xr, hr).backward()
        with torch.no_grad():
            rt = m(y.detach(), rev=True)
        res[tag] = dict(y=y.detach(), dx=x.grad, xr=xr.detach(), du=u.grad, rt=rt,
                        g={n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad})
    a, b = res["ora"], res["net"]
    for k in ("y", "dx", "xr", "du", "rt"):
        scale_ = max(1.0, a[k].abs().max().item())
        assert (a[k] - b[k]).abs().max().item() <= 2e-4 * scale_, k
    assert set(a["g"]) == set(b["g"])
    for n in a["g"]:
        ref = a["g"][n]
        tol = 2e-4 * max(ref.abs().max().item(), 1e-3)
        assert (ref - b["g"][n]).abs().max().item() <= tol, n
    assert (b["rt"] - hr).abs().max().item() < 1e-4
    assert b["y"].is_contiguous() and b["xr"].is_contiguous()


def test_standalone_freia_modules(fake_kernels):
    from sin_inn_b200.freia import modules as Fm
    Ff_o, Fm_o = R._freia()
    x = torch.randn(2, 12, 8, 12)
    sq, sq_o = Fm.IRevNetDownsampling([(12, 8, 12)]), Fm_o.IRevNetDownsampling([(12, 8, 12)])
    y = sq([x])[0]
    assert torch.equal(y, sq_o([x])[0]) and torch.equal(sq([y], rev=True)[0], x)
    assert sq.output_dims([(12, 8, 12)]) == [(48, 4, 6)]
    pm, pm_o = Fm.PermuteRandom([(12, 8, 12)], seed=3), Fm_o.PermuteRandom([(12, 8, 12)], seed=3)
    assert torch.equal(pm.perm, pm_o.perm)
    assert torch.equal(pm([x])[0], pm_o([x])[0]) and torch.equal(pm([x], rev=True)[0], pm_o([x], rev=True)[0])
    torch.manual_seed(0)
    gl = Fm.GLOWCouplingBlock([(12, 8, 12)], subnet_constructor=archs.subnet_conv, clamp=1.2)
    torch.manual_seed(0)
    gl_o = Fm_o.GLOWCouplingBlock([(12, 8, 12)], subnet_constructor=R.subnet_conv, clamp=1.2)
    gl.engine_config = None
    import os
    os.environ["SININN_PRECISION"] = "fp32"
    try:
        xg = x.clone().requires_grad_(True)
        yg = gl([xg])[0]
        yo = gl_o([x])[0]
        assert (yg - yo).abs().max() < 1e-4
        assert (gl([yg.detach()], rev=True)[0] - x).abs().max() < 1e-4
        yg.square().mean().backward()
        xo = x.clone().requires_grad_(True)
        gl_o([xo])[0].square().mean().backward()
        assert (xg.grad - xo.grad).abs().max() < 1e-5
    finally:
        del os.environ["SININN_PRECISION"]


def test_cpu_input_fails_loudly():
    opt = R.make_opt(scale=2, num_coupling=1, lr_window=1)
    net = archs.UncondSRFlow(3, 16, 16, opt)
    with pytest.raises(E.SininnError):
        net(torch.rand(1, 3, 16, 16))


@pytest.mark.parametrize("arch,scale,nc,lrw,tc", [("SRF", 4, 2, 10, False), ("IRN", 4, 1, 10, False),
                                                  # tensor-core plumbing on the stand-ins: fused 1x1 subnet, sign-bit masks,
                                                  # grouped weight gradients, coupling fused into the 3x3 conv-2 epilogue
                                                  ("SRF", 4, 2, 10, True), ("IRN", 4, 1, 10, True)])
def test_bf16_operand_path_bookkeeping(fake_kernels, arch, scale, nc, lrw, tc):
    """bf16 operand copies (cached per channel range, emitted by producer kernels) must stay coherent with
    the fp32 trunk: outputs/gradients stay within the bf16 tolerance of the fp32 oracle."""
    opt, ora, net = _pair(arch, scale, nc, lrw, 16, 32)
    net.engine_config = E.EngineConfig(precision="bf16", tensor_core=tc)
    hr, lr, z = R.synthetic_batch(opt, 2, 16, 32, seed=4)
    lrz = torch.cat((lr, z), 1)
    out = {}
    for tag, m in (("ora", ora), ("net", net)):
        for p in m.parameters():
            p.grad = None
        y = m(hr)
        R.reconstruction(y[:, :opt.lr_dims], lr).backward()
        xr = m(lrz, rev=True)
        R.reconstruction(xr, hr).backward()
        out[tag] = (y.detach(), xr.detach(), {n: p.grad for n, p in m.named_parameters() if p.requires_grad})
    for i in (0, 1):
        ref = out["ora"][i]
        assert (ref - out["net"][i]).abs().max() <= 2e-2 * max(1.0, ref.abs().max().item())
    for n, g in out["ora"][2].items():
        rel = (g - out["net"][2][n]).norm() / (g.norm() + 1e-12)
        assert rel < 1e-1, (n, float(rel))   # tiny 2x4-pixel grids: little averaging of bf16 rounding


def test_unmodified_reference_archs_runs_on_the_freia_dropin(fake_kernels):
    """INTEGRATION.md level 2: alias sin_inn_b200.freia as FrEIA and import the reference's own archs.py."""
    import importlib
    import sys
    ref = "/root/reference"
    import os
    if not os.path.exists(os.path.join(ref, "archs.py")):
        pytest.skip("reference checkout not present (GPU box)")
    import sin_inn_b200.freia as f
    saved = {k: sys.modules.get(k) for k in ("FrEIA", "FrEIA.framework", "FrEIA.modules", "archs")}
    sys.modules.update({"FrEIA": f, "FrEIA.framework": f.framework, "FrEIA.modules": f.modules})
    sys.modules.pop("archs", None)
    sys.path.insert(0, ref)
    try:
        ref_archs = importlib.import_module("archs")
        assert ref_archs.__file__.startswith(ref)
        opt = R.make_opt(scale=4, num_coupling=2, lr_window=10)
        torch.manual_seed(3)
        net = ref_archs.UncondSRFlow(3, 16, 32, opt)
        net.engine_config = E.EngineConfig(precision="fp32", tensor_core=False)
        torch.manual_seed(3)
        ora = R.build_srf(3, 16, 32, opt)
        x = torch.rand(2, 3, 16, 32)
        y = net(x)
        assert (y - ora(x)).abs().max() < 1e-4
        assert (net(y, rev=True) - x).abs().max() < 1e-4
    finally:
        sys.path.remove(ref)
        sys.modules.pop("archs", None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_direct_gradient_accumulation_matches_autograd_accumulation(fake_kernels):
    """train.SingleVideoTrainer mode: weight-gradient kernels add straight into pre-allocated param.grad."""
    opt, ora, net = _pair("SRF", 4, 2, 10, 16, 32)
    hr, lr, z = R.synthetic_batch(opt, 2, 16, 32, seed=9)
    lrz = torch.cat((lr, z), 1)
    net.plan().direct_grad = True
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    ptrs = [p.grad.data_ptr() for p in net.parameters()]
    for m in (ora, net):
        R.reconstruction(m(hr)[:, :opt.lr_dims], lr).backward()
        R.reconstruction(m(lrz, rev=True), hr).backward()
    assert ptrs == [p.grad.data_ptr() for p in net.parameters()]          # accumulated in place
    for (n, a), (_, b) in zip(ora.named_parameters(), net.named_parameters()):
        assert (a.grad - b.grad).abs().max() <= 2e-4 * max(a.grad.abs().max().item(), 1e-3), n


def test_fixed1x1conv_host(fake_kernels):
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shared_checks import fixed1x1_checks
    fixed1x1_checks("cpu")


def test_actnorm_latent_input_and_jacobian_host_logic(fake_kernels):
    """Plan ops added for SURVEY 8a7 / 8f1 on the torch stand-in kernels (CPU): an ActNorm node inside a chain
    (data-dependent init, parameter gradients through the recompute-from-inverse backward), the (lr, z) LatentInput
    entry of the inverse pass, and the coupling log-determinant -- all against the oracle shim."""
    OFf, OFm = R._freia()
    from sin_inn_b200.freia import framework as Ff, modules as Fm

    def build(Ff_, Fm_, conv):
        torch.manual_seed(4)
        nodes = [Ff_.InputNode(3, 16, 16, name="input")]
        for name, cls, args in (("sq0", Fm_.IRevNetDownsampling, {}), ("sq1", Fm_.IRevNetDownsampling, {}),
                                ("an0", Fm_.ActNorm, {}), ("g0", Fm_.GLOWCouplingBlock, {"subnet_constructor": conv, "clamp": 1.2}),
                                ("p0", Fm_.PermuteRandom, {"seed": 1})):
            nodes.append(Ff_.Node(nodes[-1], cls, args, name=name))
        nodes.append(Ff_.OutputNode(nodes[-1], name="output"))
        return Ff_.ReversibleGraphNet(nodes, verbose=False)

    ora, net = build(OFf, OFm, R.subnet_conv), build(Ff, Fm, archs.subnet_conv)
    net.engine_config = E.EngineConfig(precision="fp32")
    x = torch.rand(2, 3, 16, 16, generator=torch.Generator().manual_seed(6))
    res = {}
    for tag, m in (("ora", ora), ("net", net)):
        xi = x.clone().requires_grad_(True)
        y = m(xi)
        (y ** 2).mean().backward()
        res[tag] = (y.detach(), xi.grad, {n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad})
    (ya, dxa, ga), (yb, dxb, gb) = res["ora"], res["net"]
    assert (ya - yb).abs().max() <= 1e-4 * max(1.0, ya.abs().max()) and (dxa - dxb).abs().max() <= 1e-4 * max(1.0, dxa.abs().max())
    assert set(ga) == set(gb)
    for n in ga:
        assert (ga[n] - gb[n]).abs().max() <= 2e-4 * max(ga[n].abs().max().item(), 1e-3), n
    # LatentInput: the cat happens in the entry kernel
    lr, z = yb[:, :20].contiguous(), yb[:, 20:].contiguous()
    with torch.no_grad():
        a = net(torch.cat((lr, z), 1), rev=True)
        b = net(E.LatentInput(lr, z), rev=True)
    assert torch.equal(a, b) and (a - x).abs().max() <= 1e-5
    # coupling log-determinant
    torch.manual_seed(2)
    ob = OFm.GLOWCouplingBlock([(48, 4, 4)], subnet_constructor=R.subnet_conv_1x1, clamp=1.2)
    torch.manual_seed(2)
    nb = Fm.GLOWCouplingBlock([(48, 4, 4)], subnet_constructor=archs.subnet_conv_1x1, clamp=1.2)
    v = torch.randn(3, 48, 4, 4, generator=torch.Generator().manual_seed(1))
    import os
    os.environ["SININN_PRECISION"] = "fp32"
    try:
        for rev in (False, True):
            with torch.no_grad():
                ob([v], rev=rev)
            assert (nb.jacobian([v], rev=rev) - ob.jacobian([v], rev=rev)).abs().max() <= 1e-4 * max(1.0, ob.last_jac.abs().max())
    finally:
        del os.environ["SININN_PRECISION"]


@pytest.mark.parametrize("arch,precision,tc", [("SRF", "fp32", False), ("SRF", "bf16", True), ("SRF", "fp32tc", True),
                                               ("IRN", "fp32", False), ("IRN", "bf16", True)])
def test_stored_and_recomputed_subnet_state_agree(fake_kernels, arch, precision, tc):
    """EngineConfig.activations: "store" keeps every subnet's operand copy / hidden activation / output from the value
    pass, "recompute" re-evaluates them from the trunk the inverse restores.  Same outputs; gradients equal up to the
    round-off of the restored trunk (fp32: 1e-5; bf16 operands: their rounding decides a few copies differently)."""
    res = {}
    for mode in ("store", "recompute"):
        opt, ora, net = _pair(arch, 4, 2 if arch == "SRF" else 1, 10, 16, 32)
        net.engine_config = E.EngineConfig(precision=precision, tensor_core=tc, activations=mode)
        hr, lr, z = R.synthetic_batch(opt, 2, 16, 32, seed=4)
        x = hr.clone().requires_grad_(True)
        y = net(x)
        R.reconstruction(y[:, :opt.lr_dims], lr).backward()
        u = torch.cat((lr, z), 1).requires_grad_(True)
        R.reconstruction(net(u, rev=True), hr).backward()
        res[mode] = (y.detach(), x.grad, u.grad, {n: p.grad for n, p in net.named_parameters() if p.requires_grad})
    tol = 1e-5 if precision == "fp32" else (1e-4 if precision == "fp32tc" else 3e-2)
    a, b = res["store"], res["recompute"]
    assert torch.equal(a[0], b[0])
    for i in (1, 2):
        assert (a[i] - b[i]).norm() <= tol * max(1e-6, float(b[i].norm()))
    for n, g in a[3].items():
        assert (g - b[3][n]).norm() <= tol * max(1e-6, float(g.norm())), n
    with pytest.raises(E.SininnError):
        net.engine_config = E.EngineConfig(activations="sometimes")
        net(hr.clone().requires_grad_(True))


@pytest.mark.parametrize("arch,mode", [("SRF", "store"), ("SRF", "recompute"), ("IRN", "store")])
def test_engine_level_fusions_do_not_change_results(fake_kernels, monkeypatch, arch, mode):
    """The engine's peephole fusions -- PermuteRandom folded into the adjoining coupling kernel, the two entry squeezes + layout
    change as one kernel, a DenseBlock's five weight gradients as one problem, the fused 1x1 subnet backward -- are
    re-orderings of the same arithmetic: with the torch stand-in kernels the results are identical with and without them."""
    res = {}
    for fused in (True, False):
        for name in ("FOLD_PERM", "FUSE_SQUEEZE2", "MERGE_DENSE_WGRAD", "FUSE_1X1_BWD"):
            monkeypatch.setattr(E, name, fused)
        opt, ora, net = _pair(arch, 4, 2 if arch == "SRF" else 1, 10, 16, 32)
        net.engine_config = E.EngineConfig(precision="bf16", tensor_core=True, activations=mode)
        hr, lr, z = R.synthetic_batch(opt, 2, 16, 32, seed=5)
        x = hr.clone().requires_grad_(True)
        y = net(x)
        R.reconstruction(y[:, :opt.lr_dims], lr).backward()
        u = torch.cat((lr, z), 1).requires_grad_(True)
        xr = net(u, rev=True)
        R.reconstruction(xr, hr).backward()
        res[fused] = (y.detach(), xr.detach(), x.grad, u.grad, {n: p.grad.clone() for n, p in net.named_parameters() if p.requires_grad})
    a, b = res[True], res[False]
    for i in range(4):
        assert torch.allclose(a[i], b[i], rtol=0, atol=1e-6 * max(1.0, float(b[i].abs().max()))), i
    for n, g in a[4].items():
        assert (g - b[4][n]).norm() <= 2e-2 * max(1e-6, float(g.norm())), n      # (the fused 1x1 backward rounds dh to bf16 like the kernel)
