"""Accuracy of the fp32 paths (tensor-core split operands vs CUDA cores) against the oracle evaluated in fp64 and in fp32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import ref_torch as R
import test_gpu_parity as T
DEV = "cuda:0"


def run(m, dev, dt, opt, hr, lr, z):
    for p in m.parameters():
        p.grad = None
    x = hr.to(dev, dt).clone().requires_grad_(True)
    y = m(x)
    (R.reconstruction(y[:, :opt.lr_dims], lr.to(dev, dt)) + 0.3 * R.latent_nll(y[:, opt.lr_dims:])).backward()
    u = torch.cat((lr, z), 1).to(dev, dt).clone().requires_grad_(True)
    xr = m(u, rev=True)
    R.reconstruction(xr, hr.to(dev, dt)).backward()
    with torch.no_grad():
        rt = m(y.detach(), rev=True)
    return dict(y=y.detach().cpu().double(), dx=x.grad.cpu().double(), xr=xr.detach().cpu().double(), du=u.grad.cpu().double(),
                rt=rt.cpu().double(), g={n: p.grad.detach().cpu().double() for n, p in m.named_parameters() if p.requires_grad})


def cmp(a, b, hr):
    line = [f"{k} {float((a[k] - b[k]).abs().max() / a[k].abs().max()):.1e}" for k in ("y", "xr", "dx", "du")]
    worst = max((float((a['g'][n] - b['g'][n]).abs().max() / max(a['g'][n].abs().max().item(), 1e-3)), n) for n in a['g'])
    wl2 = max(float((a['g'][n] - b['g'][n]).norm() / a['g'][n].norm()) for n in a['g'])
    return " ".join(line) + f" | wgrad worst max-rel {worst[0]:.1e} ({worst[1].split('module_list.')[-1]}) worst rel_l2 {wl2:.1e} | rt {float((b['rt'] - hr.double()).abs().max()):.1e}"


for case in T.FP32_CASES[:2] + [("SRF", 4, 4, 10, 2, 64, 64), ("SRF", 4, 4, 10, 2, 256, 256)]:
    arch, scale, nc, lrw, B, H, W = case
    opt, ora, _ = T.build_pair(arch, scale, nc, lrw, H, W, "fp32")
    hr, lr, z = R.synthetic_batch(opt, B, H, W, seed=3)
    o32 = run(ora, "cpu", torch.float32, opt, hr, lr, z)
    o64 = run(ora.double(), "cpu", torch.float64, opt, hr, lr, z)
    print(case, "\n   oracle fp32 vs fp64:", cmp(o64, o32, hr), flush=True)
    for tc in (True, False):
        _, _, net = T.build_pair(arch, scale, nc, lrw, H, W, "fp32", tensor_core=tc)
        n = run(net, DEV, torch.float32, opt, hr, lr, z)
        print("   ", "tensor cores" if tc else "CUDA cores  ", "vs fp64:", cmp(o64, n, hr), flush=True)
