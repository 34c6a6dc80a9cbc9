"""The oracle (oracle/ref_torch.py + the FrEIA shim) against the golden vectors that oracle/make_golden.py
produced by running the UNMODIFIED reference archs.py.  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_torch as R

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")) if "known" not in p)


def load_case(name):
    f = np.load(os.path.join(GOLD, name + ".npz"))
    scale, nc, lrw, B, H, W, wseed, dseed = (int(v) for v in f["meta"])
    arch = str(f["arch"])
    opt = R.make_opt(scale=scale, num_coupling=nc, lr_window=lrw, architecture=arch)
    return f, arch, opt, (B, H, W), wseed


def build_seeded(arch, opt, H, W, wseed, builder=None):
    torch.manual_seed(wseed)
    net = (builder or R.build)(arch, 3, H, W, opt)
    if arch == "IRN":
        R.randomize_irn_conv5(net, seed=1)
    return net


def stats(named):
    return np.array([[float(t.double().sum()), float(t.double().pow(2).sum().sqrt())] for _, t in named])


def test_fixture_set_is_complete():
    assert set(CASES) >= {"srf_s2_c4", "srf_s4_c2", "irn_s2_c2", "irn_s4_c1"}


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_outputs(name):
    f, arch, opt, (B, H, W), wseed = load_case(name)
    net = build_seeded(arch, opt, H, W, wseed)
    trainable = [(n, p.detach()) for n, p in net.named_parameters() if p.requires_grad]
    assert [n for n, _ in trainable] == [str(s) for s in f["param_names"]]
    np.testing.assert_allclose(stats(trainable), f["wstats"], rtol=1e-6, atol=1e-7)   # same seeded init as the reference
    hr, lr, z = (torch.from_numpy(f[k]) for k in ("hr", "lr", "z"))
    x = hr.clone().requires_grad_(True)
    y = net(x)
    (R.reconstruction(y[:, :opt.lr_dims], lr) + 0.5 * R.latent_nll(y[:, opt.lr_dims:])).backward()
    u = torch.cat((lr, z), 1).requires_grad_(True)
    xr = net(u, rev=True)
    R.reconstruction(xr, hr).backward()
    rt = net(y.detach(), rev=True)
    for k, v in (("y", y), ("dx", x.grad), ("xr", xr), ("du", u.grad), ("rt", rt)):
        ref = f[k]
        np.testing.assert_allclose(v.detach().numpy(), ref, rtol=0, atol=2e-6 * max(1.0, np.abs(ref).max()), err_msg=k)
    gstats = stats([(n, p.grad) for n, p in net.named_parameters() if p.requires_grad])
    np.testing.assert_allclose(gstats, f["gstats"], rtol=2e-4, atol=1e-7)
    named = dict(net.named_parameters())
    for i, n in enumerate(f["kept_grad_names"]):
        g = named[str(n)].grad
        g = g[:8] if g.dim() == 4 else g
        ref = f[f"kept_grad_{i}"]
        np.testing.assert_allclose(g.numpy(), ref, rtol=0, atol=1e-5 * max(np.abs(ref).max(), 1e-3))
    assert float((rt - hr).abs().max()) < 1e-5        # north-star round-trip bar


def test_known_answers():
    f = np.load(os.path.join(GOLD, "known_answers.npz"))
    Ff, Fm = R._freia()
    t = torch.from_numpy(f["squeeze_in"])
    sq = Fm.IRevNetDownsampling([tuple(t.shape[1:])])
    out = sq([t])[0]
    assert np.array_equal(out.numpy(), f["squeeze_out"])
    assert torch.equal(sq([out], rev=True)[0], t)
    # index formula out[b,(dy*2+dx)*C+c,i,j] = in[b,c,2i+dy,2j+dx]
    C = t.shape[1]
    for dy in (0, 1):
        for dx in (0, 1):
            for c in range(C):
                assert torch.equal(out[:, (dy * 2 + dx) * C + c], t[:, c, dy::2, dx::2])
    haar = R.HaarDownsampling(3)
    hx = torch.from_numpy(f["haar_in"])
    np.testing.assert_allclose(haar(hx).numpy(), f["haar_out"], atol=1e-6)
    np.testing.assert_allclose(haar(haar(hx), rev=True).numpy(), f["haar_rt"], atol=1e-6)
    for key in f.files:
        if key.startswith("perm_"):
            seed, C = int(key.split("_")[1][1:]), int(key.split("_")[2][1:])
            assert np.array_equal(Fm.PermuteRandom([(C, 1, 1)], seed=seed).perm.numpy(), f[key])
    assert f["perm_s0_c48"][:12].tolist() == [29, 4, 26, 30, 32, 37, 34, 40, 7, 10, 11, 31]
    assert f["perm_s0_c192"][:12].tolist() == [110, 74, 163, 97, 126, 71, 18, 157, 145, 7, 5, 139]


@pytest.mark.parametrize("arch,scale,nc,lrw,count", [("SRF", 2, 4, 1, 739712), ("SRF", 4, 4, 10, 3692416),
                                                     ("SRF", 4, 8, 10, 7384832), ("IRN", 2, 4, 1, 1375632),
                                                     ("IRN", 4, 4, 10, 5691408)])
def test_param_counts(arch, scale, nc, lrw, count):
    opt = R.make_opt(scale=scale, num_coupling=nc, lr_window=lrw, architecture=arch)
    net = R.build(arch, 3, 64, 64, opt)
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == count
    y = net(torch.rand(1, 3, 32, 32))
    assert y.shape == (1, 3 * 4 ** (1 + R.n_levels(scale)), 32 // 2 ** (1 + R.n_levels(scale)), 32 // 2 ** (1 + R.n_levels(scale)))


def test_irn_is_identity_haar_at_init():
    """conv5 = 0 (archs.py:86) makes every InvBlockExp the identity: the net is three nested Haars."""
    opt = R.make_opt(scale=4, num_coupling=2, lr_window=10, architecture="IRN")
    net = R.build_irn(3, 32, 32, opt)
    x = torch.rand(1, 3, 32, 32)
    y = x
    for c in (3, 12, 48):
        y = R.HaarDownsampling(c)(y)
    assert torch.allclose(net(x), y, atol=1e-6)
