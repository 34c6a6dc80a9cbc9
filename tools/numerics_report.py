"""Prints the error of the CUDA paths against the fp32 CPU oracle (and the oracle's own fp64 run) for
outputs, input gradients and weight gradients: relative Frobenius error and max error / max|ref|."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R
from sin_inn_b200 import archs

def run(arch, scale, nc, lrw, B, H, W, precision, tc, seed=11):
    opt = R.make_opt(scale=scale, num_coupling=nc, lr_window=lrw, architecture=arch, precision=precision, tensor_core=tc)
    torch.manual_seed(seed); ora = R.build(arch, 3, H, W, opt)
    torch.manual_seed(seed); net = {"SRF": archs.UncondSRFlow, "IRN": archs.InvRescaleNet}[arch](3, H, W, opt)
    if arch == "IRN":
        R.randomize_irn_conv5(ora, 1); R.randomize_irn_conv5(net, 1)
    ora = ora.double(); net = net.cuda()
    hr, lr, z = R.synthetic_batch(opt, B, H, W, seed=3)
    lrz = torch.cat((lr, z), 1)
    res = {}
    for tag, m, dev, dt in (("ora", ora, "cpu", torch.float64), ("net", net, "cuda", torch.float32)):
        x = hr.to(dev, dt).requires_grad_(True)
        y = m(x)
        R.reconstruction(y[:, :opt.lr_dims], lr.to(dev, dt)).backward()
        u = lrz.to(dev, dt).requires_grad_(True)
        xr = m(u, rev=True)
        R.reconstruction(xr, hr.to(dev, dt)).backward()
        with torch.no_grad():
            rt = m(y.detach(), rev=True)
        res[tag] = dict(y=y, dx=x.grad, xr=xr, du=u.grad, rt=rt, **{"g:" + n: p.grad for n, p in m.named_parameters() if p.requires_grad})
    rows = {}
    for k, ref in res["ora"].items():
        if k == "rt":
            continue
        got = res["net"][k].detach().double().cpu(); ref = ref.detach().double()
        rows[k] = (float((ref - got).norm() / ref.norm().clamp_min(1e-30)), float((ref - got).abs().max() / ref.abs().max().clamp_min(1e-30)))
    gk = [k for k in rows if k.startswith("g:")]
    worst_l2 = max(gk, key=lambda k: rows[k][0]); worst_mx = max(gk, key=lambda k: rows[k][1])
    rt_err = (res["net"]["rt"].cpu().double() - hr.double()).abs()
    out = dict(case=f"{arch} s{scale} c{nc} B{B} {H}x{W} {precision} tc={tc}",
               **{k: [f"{rows[k][0]:.2e}", f"{rows[k][1]:.2e}"] for k in ("y", "xr", "dx", "du")},
               wgrad_rel_l2_median=f"{sorted(rows[k][0] for k in gk)[len(gk)//2]:.2e}",
               wgrad_rel_l2_worst=[worst_l2, f"{rows[worst_l2][0]:.2e}"], wgrad_max_worst=[worst_mx, f"{rows[worst_mx][1]:.2e}"],
               roundtrip_max=f"{float(rt_err.max()):.2e}", roundtrip_mean=f"{float(rt_err.mean()):.2e}")
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for prec, tc in (("fp32", False), ("bf16", False), ("bf16", True)):
        run("SRF", 4, 4, 10, 2, 128, 128, prec, tc)
        run("IRN", 4, 2, 10, 2, 64, 64, prec, tc)
