"""InvRescaleNet (IRN, archs.py:201-233) training step at the bench shape: eager vs CUDA-graph time per step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R
from sin_inn_b200 import archs, kernels, train

dev = torch.device("cuda", 0)
P, B = int(os.environ.get("P", 256)), int(os.environ.get("B", 32))
opt = R.make_opt(scale=4, num_coupling=4, lr_window=10, architecture="IRN", precision="bf16")
torch.manual_seed(0)
net = archs.InvRescaleNet(3, P, P, opt)
R.randomize_irn_conv5(net, 1)
tr = train.SingleVideoTrainer(net.to(dev), opt)
batch = tuple(t.to(dev) for t in R.synthetic_batch(opt, B, P, P, seed=0))
kernels.LAUNCHES = 0
tr.training_step(*batch)
n = kernels.LAUNCHES
step = tr.capture(*batch, warmup=2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step(*batch)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
# algorithmic FLOPs: 20.18 GFLOP per 256^2 patch and direction, x 6 for a training step
print(f"IRN scale4 c4 train step, batch {B} x {P}x{P}: {ms:.3f} ms/step, {B / ms * 1e3:.0f} patches/s, {n} launches/step, "
      f"{6 * 20.18e9 * (P / 256) ** 2 * B / (ms / 1e3) / 1e12:.0f} algorithmic TFLOP/s")
