"""Data-parallel host logic with world_size 2 on CPU (gloo): sharding helpers, the flat gradient arena,
one all-reduce per step, identical parameters on every rank afterwards.  The INN itself is replaced by a tiny
torch module here (the kernels need a GPU); the code path under test is sin_inn_b200/train.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_kernels as FK
    from sin_inn_b200 import train
    train.K = FK                                   # CPU stand-in for the fused Adam kernel
    train.engine.invalidate_packs = lambda: None
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 3, 3, padding=1))
    flat = train.FlatParams(net)
    opt = train.FusedAdam(flat, lr=1e-2, betas=(0.9, 0.99), weight_decay=1e-5)
    if world > 1:
        dist.broadcast(flat.flat, src=0)
    g = torch.Generator().manual_seed(100 + rank)          # per-rank data shard
    for _ in range(3):
        opt.zero_grad()
        x = torch.rand(4, 3, 8, 8, generator=g)
        net(x).square().mean().backward()
        (net(x * 0.5) - x).square().mean().backward()       # two backward passes accumulate, as in the train step
        if world > 1:
            dist.all_reduce(flat.grad)
        opt.step(grad_scale=1.0 / world)
    out[rank] = flat.flat.clone()
    frames = list(train.shard_frames(120, rank, world))
    out[f"frames{rank}"] = (frames[0], frames[-1], len(frames))
    dist.destroy_process_group()


def _run(world, port):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    return dict(out)


def test_two_ranks_stay_in_sync_and_match_large_batch():
    res = _run(2, 29611)
    assert torch.equal(res[0], res[1])                      # replicas identical after all-reduce + fused Adam
    assert res["frames0"] == (0, 59, 60) and res["frames1"] == (60, 119, 60)


def test_flat_arena_keeps_module_semantics():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from sin_inn_b200 import train
    net = torch.nn.Conv2d(2, 2, 1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    flat = train.FlatParams(net)
    assert flat.numel == sum(p.numel() for p in net.parameters())
    for k, v in net.state_dict().items():
        assert torch.equal(v, sd[k])
    net(torch.rand(1, 2, 4, 4)).sum().backward()
    assert net.weight.grad.data_ptr() == flat.grad.data_ptr()          # gradients land in the arena
    flat.zero_grad()
    assert float(flat.grad.abs().sum()) == 0.0
    net.load_state_dict(sd)                                            # checkpoints still load through the views
    assert torch.equal(flat.flat[:net.weight.numel()].view_as(net.weight), sd["weight"])
