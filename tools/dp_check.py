"""Two ranks, each with half of a batch, must end up with the parameters one rank gets from the whole batch (the
gradients are means over the batch, so the all-reduced sum / world equals the large-batch gradient up to fp32
summation order).  Run with torch.distributed.run --nproc-per-node 2."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import archs, train, config as Cg

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
opt = Cg.make_opt(scale=4, num_coupling=2, lr_window=10, precision="fp32")
hr, lr, z = Cg.synthetic_batch(opt, 8, 64, 64, seed=5)


def make(ws):
    torch.manual_seed(0)
    return train.SingleVideoTrainer(archs.UncondSRFlow(3, 64, 64, opt).to(dev), opt, world_size=ws)


dp = make(world)
dp.broadcast_params()
per = 8 // world
sl = slice(rank * per, (rank + 1) * per)
for _ in range(3):
    dp.training_step(hr[sl].to(dev), lr[sl].to(dev), z[sl].to(dev))
torch.cuda.synchronize()
if rank == 0:
    one = make(1)
    for _ in range(3):
        one.training_step(hr.to(dev), lr.to(dev), z.to(dev))
    torch.cuda.synchronize()
    d = (dp.flat.flat - one.flat.flat).abs().max().item()
    upd = (one.flat.flat - make(1).flat.flat).abs().max().item()
    print(f"max |param(2 ranks x {per}) - param(1 rank x 8)| after 3 steps: {d:.3e} (largest update {upd:.3e})")
    assert d <= 2e-2 * upd, "data-parallel step diverges from the large-batch step"
    print("DP check ok")
other = [torch.zeros_like(dp.flat.flat) for _ in range(world)]
dist.all_gather(other, dp.flat.flat)
assert all(torch.equal(o, other[0]) for o in other), "ranks diverged"
if rank == 0:
    print("replicas bit-identical")
dist.destroy_process_group()
