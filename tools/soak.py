"""Soak test: many graph-replayed training steps on two alternating synthetic batches; losses stay finite and fall,
device memory stays flat."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R
from sin_inn_b200 import archs, train

dev = torch.device("cuda", 0)
N = int(os.environ.get("STEPS", 600))
opt = R.make_opt(scale=4, num_coupling=4, lr_window=10, precision="bf16")
torch.manual_seed(0)
tr = train.SingleVideoTrainer(archs.UncondSRFlow(3, 256, 256, opt).to(dev), opt)
batches = [tuple(t.to(dev) for t in R.synthetic_batch(opt, 32, 256, 256, seed=s)) for s in range(2)]
step = tr.capture(*batches[0], warmup=3)
hist, mem0 = [], None
for i in range(N):
    lf, lb = step(*batches[i % 2])
    if i % 100 == 0 or i == N - 1:
        torch.cuda.synchronize()
        hist.append((i, float(lf), float(lb), torch.cuda.memory_allocated() / 2**20))
        if mem0 is None:
            mem0 = hist[-1][3]
for h in hist:
    print("step %4d  fwd_loss %.5f  bwd_loss %.5f  allocated %.0f MiB" % h)
assert all(torch.isfinite(torch.tensor(h[1:3])).all() for h in hist)
assert hist[-1][2] < hist[0][2] and hist[-1][1] < hist[0][1], "losses did not fall"
assert abs(hist[-1][3] - mem0) < 1.0, "device memory grew"
assert torch.isfinite(tr.flat.flat).all()
print("soak ok")
