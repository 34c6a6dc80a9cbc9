"""ORACLE / TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the UNMODIFIED reference
/root/reference/archs.py (UncondSRFlow archs.py:19-71 through the FrEIA shim in
oracle/freia_shim; InvRescaleNet archs.py:201-233 directly) on seeded inputs,
and checks on the way that oracle/ref_torch.py reproduces it bit-for-bit in
fp32 with the same seed (same RNG consumption => same weights) and to 1e-6 with
the reference's state_dict loaded.

Run here only (needs /root/reference):  python oracle/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

CASES = [
    # name, arch, scale, num_coupling, lr_window, batch, H, W
    ("srf_s2_c4", "SRF", 2, 4, 1, 2, 32, 32),     # BASELINE.json configs[0] shape family (scale 2, lr_dims 12)
    ("srf_s4_c2", "SRF", 4, 2, 10, 2, 32, 48),    # two levels, default lr_dims 84, non-square
    ("irn_s2_c2", "IRN", 2, 2, 1, 2, 32, 32),
    ("irn_s4_c1", "IRN", 4, 1, 10, 1, 32, 48),    # 84/108 uneven split on the second level
]
WSEED = 1234


def import_reference_archs():
    shim = os.path.join(ROOT, "oracle", "freia_shim")
    sys.path.insert(0, shim)
    sys.path.insert(1, REF)
    try:
        import archs  # the reference's own file, unmodified
    finally:
        sys.path.remove(shim)
        sys.path.remove(REF)
    assert os.path.abspath(archs.__file__).startswith(REF)
    return archs


def param_stats(named):
    return np.array([[float(t.double().sum()), float(t.double().pow(2).sum().sqrt())] for _, t in named],
                    dtype=np.float64)


def run_case(archs, R, name, arch, scale, nc, lr_window, B, H, W):
    opt = R.make_opt(scale=scale, num_coupling=nc, lr_window=lr_window, architecture=arch)
    ctor = {"SRF": archs.UncondSRFlow, "IRN": archs.InvRescaleNet}[arch]
    torch.manual_seed(WSEED)
    ref = ctor(3, H, W, opt)
    if arch == "IRN":
        R.randomize_irn_conv5(ref, seed=1)
    torch.manual_seed(WSEED)
    ora = R.build(arch, 3, H, W, opt)
    if arch == "IRN":
        R.randomize_irn_conv5(ora, seed=1)
    sd_ref, sd_ora = ref.state_dict(), ora.state_dict()
    assert list(sd_ref) == list(sd_ora), "state_dict keys differ"
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd_ora[k]), f"seeded init differs at {k}"
    ora.load_state_dict(sd_ref)  # the reference's own checkpoint format must load

    hr, lr, z = R.synthetic_batch(opt, B, H, W, seed=7)
    lrz = torch.cat((lr, z), 1)
    out = {}
    for tag, net in (("ref", ref), ("ora", ora)):
        for p in net.parameters():
            p.grad = None
        x = hr.clone().requires_grad_(True)
        y = net(x)
        loss_f = R.reconstruction(y[:, :opt.lr_dims], lr) + 0.5 * R.latent_nll(y[:, opt.lr_dims:])
        loss_f.backward()
        u = lrz.clone().requires_grad_(True)
        xr = net(u, rev=True)
        loss_b = R.reconstruction(xr, hr)
        loss_b.backward()
        rt = net(y.detach(), rev=True)
        named = [(n, p.grad) for n, p in net.named_parameters() if p.requires_grad]
        out[tag] = dict(y=y.detach(), dx=x.grad, xr=xr.detach(), du=u.grad, rt=rt.detach(),
                        gstats=param_stats(named), loss=np.array([float(loss_f.detach()), float(loss_b.detach())]),
                        grads={n: g.clone() for n, g in named})
    for k in ("y", "dx", "xr", "du", "rt"):
        d = (out["ref"][k] - out["ora"][k]).abs().max().item()
        s = out["ref"][k].abs().max().item()
        assert d <= 2e-6 * max(1.0, s), f"{name}: oracle deviates from reference on {k}: {d} (scale {s})"
    for n in out["ref"]["grads"]:
        a, b = out["ref"]["grads"][n], out["ora"]["grads"][n]
        d = (a - b).abs().max().item()
        assert d <= 1e-5 * max(a.abs().max().item(), 1e-3), f"{name}: grad {n} deviates {d}"
    r = out["ref"]
    trainable = [(n, p.detach()) for n, p in ref.named_parameters() if p.requires_grad]
    # a few full weight-gradient tensors (first/last trainable) + stats for all
    names = [n for n, _ in trainable]
    keep = [names[0], names[1], names[-2], names[-1]]
    fix = dict(
        meta=np.array([scale, nc, lr_window, B, H, W, WSEED, 7], dtype=np.int64),
        arch=np.array(arch), param_names=np.array(names),
        hr=hr.numpy(), lr=lr.numpy(), z=z.numpy(),
        y=r["y"].numpy(), dx=r["dx"].numpy(), xr=r["xr"].numpy(), du=r["du"].numpy(), rt=r["rt"].numpy(),
        loss=r["loss"], wstats=param_stats(trainable), gstats=r["gstats"],
        kept_grad_names=np.array(keep),
    )
    for i, n in enumerate(keep):
        g = r["grads"][n]
        fix[f"kept_grad_{i}"] = (g[:8] if g.dim() == 4 else g).numpy()   # weights: first 8 output channels
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **fix)
    print(f"{name}: wrote {path} ({os.path.getsize(path)/1024:.0f} KiB); "
          f"round trip {float((r['rt']-hr).abs().max()):.2e}; losses {r['loss']}")


def known_answers(archs, R):
    """Small known-answer vectors that do not depend on weights."""
    Ff, Fm = R._freia()
    t = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).reshape(2, 3, 4, 6)
    sq = Fm.IRevNetDownsampling([(3, 4, 6)])([t])[0]
    haar_ref = archs.HaarDownsampling(3)
    g = torch.Generator().manual_seed(3)
    hx = torch.randn(2, 3, 6, 8, generator=g)
    hy = haar_ref(hx)
    hxr = haar_ref(hy, rev=True)
    hy2 = R.HaarDownsampling(3)(hx)
    assert (hy - hy2).abs().max() < 1e-6
    perms = {}
    for seed, C in ((0, 48), (1, 48), (2, 48), (3, 48), (0, 192), (1, 192), (0, 12)):
        perms[f"perm_s{seed}_c{C}"] = Fm.PermuteRandom([(C, 1, 1)], seed=seed).perm.numpy()
    path = os.path.join(ROOT, "tests", "golden", "known_answers.npz")
    np.savez_compressed(path, squeeze_in=t.numpy(), squeeze_out=sq.numpy(),
                        haar_in=hx.numpy(), haar_out=hy.numpy(), haar_rt=hxr.numpy(), **perms)
    print("known answers ->", path)


def main():
    archs = import_reference_archs()
    from oracle import ref_torch as R
    torch.set_num_threads(8)
    known_answers(archs, R)
    for case in CASES:
        run_case(archs, R, *case)


if __name__ == "__main__":
    main()
