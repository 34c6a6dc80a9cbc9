"""Debug aid: run the same net through the real kernels (GPU) and the torch stand-ins (CPU) and report the
first executed op whose trunk differs."""
import os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_torch as R
from sin_inn_b200 import archs, engine as E
import fake_kernels as FK

def run(arch, scale, nc, lrw, B, H, W, precision, rev=False, seed=1234):
    opt = R.make_opt(scale=scale, num_coupling=nc, lr_window=lrw, architecture=arch, precision=precision, tensor_core=False)
    torch.manual_seed(seed)
    net = {"SRF": archs.UncondSRFlow, "IRN": archs.InvRescaleNet}[arch](3, H, W, opt)
    if arch == "IRN":
        R.randomize_irn_conv5(net, 1)
    hr, lr, z = R.synthetic_batch(opt, B, H, W, seed=7)
    x = torch.cat((lr, z), 1) if rev else hr
    realK, real_req = E.K, E.require_cuda
    # CPU stand-in pass
    E.K, E.require_cuda = FK, (lambda t, what="tensor": None)
    E._pack_cache.clear(); E.TRACE = []
    with torch.no_grad():
        y_cpu = net(x, rev=rev)
    t_cpu = E.TRACE
    # GPU pass
    E.K, E.require_cuda = realK, real_req
    E._pack_cache.clear(); E.TRACE = []
    net = net.to("cuda")
    with torch.no_grad():
        y_gpu = net(x.cuda(), rev=rev)
    t_gpu = E.TRACE; E.TRACE = None
    print(f"== {arch} s{scale} c{nc} B{B} {H}x{W} {precision} rev={rev}: final err {(y_cpu - y_gpu.cpu()).abs().max().item():.3e}")
    for (la, a), (lb, b) in zip(t_cpu, t_gpu):
        err = (a - b).abs().max().item()
        print(f"   {la:40s} shape {tuple(a.shape)} err {err:.3e}")

if __name__ == "__main__":
    run("SRF", 2, 4, 1, 4, 64, 64, "bf16")
    run("SRF", 2, 4, 1, 4, 64, 64, "fp32")
    run("SRF", 4, 2, 10, 2, 32, 48, "fp32")
    run("IRN", 4, 1, 10, 1, 32, 48, "fp32")
    run("SRF", 4, 2, 10, 2, 32, 48, "fp32", rev=True)
