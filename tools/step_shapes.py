"""Per-SHAPE timing of one training step at the bench workload: the eager single-stream step with a CUDA-event pair
around every launch (as bench.py's family pass does), grouped by kernel family AND launch shape.
  python tools/step_shapes.py            (env: ARCH=SRF|IRN, B, P, PRECISION, ACT=auto|store|recompute, STEPS)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import archs, kernels, train
from sin_inn_b200 import config as R

dev = torch.device("cuda", 0)
P, B = int(os.environ.get("P", 256)), int(os.environ.get("B", 32))
STEPS = int(os.environ.get("STEPS", 5))
ARCH = os.environ.get("ARCH", "SRF")
opt = R.make_opt(scale=4, num_coupling=4, lr_window=10, precision=os.environ.get("PRECISION", "bf16"),
                 activations=os.environ.get("ACT", "auto"), architecture=ARCH)
torch.manual_seed(0)
net = (archs.UncondSRFlow if ARCH == "SRF" else archs.InvRescaleNet)(3, P, P, opt).to(dev)
if ARCH != "SRF":            # conv5 is zero-initialised in the reference: a fresh IRN is the identity
    g = torch.Generator(device="cpu").manual_seed(1)
    for m in net.modules():
        if isinstance(m, archs.DenseBlock):
            m.conv5.weight.data.copy_(0.02 * torch.randn(m.conv5.weight.shape, generator=g))
tr = train.SingleVideoTrainer(net, opt)
tr.overlap = False
net.plan().side_wgrad = False
batch = tuple(t.to(dev) for t in R.synthetic_batch(opt, B, P, P, seed=0, with_z=False)[:2])
for _ in range(3):
    tr.training_step(*batch)
torch.cuda.synchronize()
kernels.profile_begin()
for _ in range(STEPS):
    torch.cuda._sleep(int(0.04 * 1.9e9))
    tr.training_step(*batch)
    torch.cuda.synchronize()
prof = kernels.profile_end(by_shape=True)
tot = sum(v["ms"] for v in prof.values())
print(f"{ARCH} B={B} P={P}: eager single-stream step {tot / STEPS:.3f} ms (sum of per-launch event times)")
print("| family shape | launches/step | ms/step | share | mean us | TFLOP/s | GB/s |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    n = v["n"] / STEPS
    tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0
    gb = v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["bytes"] else 0
    print(f"| {k} | {n:.0f} | {v['ms'] / STEPS:.3f} | {100 * v['ms'] / tot:.1f}% | {1e3 * v['ms'] / v['n']:.1f} | {tf:.0f} | {gb:.0f} |")
