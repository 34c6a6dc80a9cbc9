"""Per-shape timing of 1080p inference (forward HR -> (LR, z), inverse (LR, z) -> HR) at the bench micro-batch: eager launches
with a CUDA-event pair each, grouped by kernel family and launch shape.  env: MB (micro-batch, default 2), REPS."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import archs, kernels
from sin_inn_b200 import config as R

dev = torch.device("cuda", 0)
MB, REPS = int(os.environ.get("MB", 2)), int(os.environ.get("REPS", 4))
opt = R.make_opt(scale=4, num_coupling=4, lr_window=10, precision="bf16")
torch.manual_seed(0)
net = archs.UncondSRFlow(3, 1080, 1920, opt).to(dev).eval()
hr = torch.rand(MB, 3, 1080, 1920, device=dev)
with torch.no_grad():
    lrz = net(hr)
    for direction, fn in (("forward", lambda: net(hr)), ("inverse", lambda: net(lrz, rev=True))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        kernels.profile_begin()
        for _ in range(REPS):
            torch.cuda._sleep(int(0.02 * 1.9e9))
            fn()
            torch.cuda.synchronize()
        prof = kernels.profile_end(by_shape=True)
        tot = sum(v["ms"] for v in prof.values())
        print(f"{direction}: {tot / REPS:.3f} ms per micro-batch of {MB} frames ({MB * REPS / (tot / 1e3):.0f} frames/s, sum of per-launch event times)")
        print("| family shape | launches | ms | share | mean us | TFLOP/s | GB/s |")
        print("|---|---:|---:|---:|---:|---:|---:|")
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0
            gb = v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["bytes"] else 0
            print(f"| {k} | {v['n'] / REPS:.0f} | {v['ms'] / REPS:.3f} | {100 * v['ms'] / tot:.1f}% | {1e3 * v['ms'] / v['n']:.1f} | {tf:.0f} | {gb:.0f} |")
