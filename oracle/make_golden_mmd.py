"""ORACLE TOOLING (run in the build container only; /root/reference does not exist on the GPU box).
Golden values of loss.mmd (reference loss.py:9-36) on seeded inputs.  The reference function cannot run on CPU as
written (it allocates its accumulators with .to('cuda'), loss.py:27-29); this script executes the reference SOURCE
with exactly that token removed -- nothing else changes -- in float64 and float32, and stores inputs + values +
autograd gradients in tests/golden/mmd_known.npz.  It also checks oracle/ref_torch.mmd against it."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R  # noqa: E402

src = open("/root/reference/loss.py").read()
assert src.count(".to('cuda')") == 3
ref = types.ModuleType("reference_loss")
exec(compile(src.replace(".to('cuda')", ""), "/root/reference/loss.py", "exec"), ref.__dict__)

out = {}
torch.manual_seed(123)
for tag, shape in (("a", (4, 6, 5, 7)), ("b", (8, 3, 16, 16))):
    x = torch.rand(*shape, dtype=torch.float64)
    y = torch.rand(*shape, dtype=torch.float64) * 0.8 + 0.1
    out[f"{tag}_x"], out[f"{tag}_y"] = x.float().numpy(), y.float().numpy()
    for rev in (False, True):
        xs = x.float().double().clone().requires_grad_(True)       # the stored fp32 inputs, evaluated in fp64
        v = ref.mmd(xs, y.float().double(), rev)
        v.backward()
        mine = R.mmd(x.float().double(), y.float().double(), rev)
        assert abs(float(v) - float(mine)) <= 1e-6 * abs(float(v)), (float(v), float(mine))
        out[f"{tag}_val_{int(rev)}"] = np.float64(float(v))
        out[f"{tag}_grad_{int(rev)}"] = xs.grad.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mmd_known.npz"), **out)
print("wrote tests/golden/mmd_known.npz", {k: float(v) for k, v in out.items() if "val" in k})
