#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== gpu tests"; python -m pytest tests -q -x -m gpu -k "wgrad or IRN or irn" 2>&1 | tail -4
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== irn profile"; python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from sin_inn_b200 import archs, train, kernels, config as R
dev = torch.device('cuda')
opt = R.make_opt(scale=4, num_coupling=4, lr_window=10, precision='bf16', architecture='IRN')
torch.manual_seed(0)
net = archs.InvRescaleNet(3, 256, 256, opt).to(dev)
g = torch.Generator().manual_seed(1)
for m in net.modules():
    if isinstance(m, archs.DenseBlock):
        m.conv5.weight.data.copy_(0.02 * torch.randn(m.conv5.weight.shape, generator=g))
tr = train.SingleVideoTrainer(net, opt)
tr.overlap = False; net.plan().side_wgrad = False
hr, lr, _ = (t.to(dev) if t is not None else None for t in R.synthetic_batch(opt, 32, 256, 256, with_z=False))
for _ in range(2): tr.training_step(hr, lr)
kernels.LAUNCHES = 0
kernels.profile_begin()
tr.training_step(hr, lr)
p = kernels.profile_end()
print("launches", kernels.LAUNCHES, {k: round(v['ms'], 3) for k, v in sorted(p.items(), key=lambda kv: -kv[1]['ms'])})
PY
} > gpurun_out/r2m.log 2>&1
tail -30 gpurun_out/r2m.log
