"""Eager vs CUDA-graph training step at the bench workload: GPU time per step, host enqueue time per step,
and a check that the replayed graph produces the same parameters as the eager step."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R
from sin_inn_b200 import archs, kernels, train

dev = torch.device("cuda", 0)
P, B = int(os.environ.get("P", 256)), int(os.environ.get("B", 32))
opt = R.make_opt(scale=4, num_coupling=4, lr_window=10, precision="bf16")


def make():
    torch.manual_seed(0)
    net = archs.UncondSRFlow(3, P, P, opt).to(dev)
    return train.SingleVideoTrainer(net, opt)


batches = [tuple(t.to(dev) for t in R.synthetic_batch(opt, B, P, P, seed=i)) for i in range(2)]


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1e3 * host / n


ta = make()
for i in range(3):
    ta.training_step(*batches[i % 2])
kernels.LAUNCHES = 0
gpu, host = timed(lambda i: ta.training_step(*batches[i % 2]), 10)
print(f"eager : {gpu:.3f} ms/step on the GPU, {host:.3f} ms/step host enqueue, {kernels.LAUNCHES // 10} launches/step")

tb = make()
step = tb.capture(*batches[0], warmup=3)
gpu, host = timed(lambda i: step(*batches[i % 2]), 10)
print(f"graph : {gpu:.3f} ms/step on the GPU, {host:.3f} ms/step host")
# both trainers have now done 13 steps on the same batch sequence?  eager: 3 + 10 alternating 0,1,0,...; graph: 3 warm-up
# steps on batch 0 then 10 alternating -> not the same sequence; compare a fresh pair instead
tc, td = make(), make()
stepd = td.capture(*batches[0], warmup=0) if False else None
for i in range(3):
    tc.training_step(*batches[0])
stepd = td.capture(*batches[0], warmup=3)
for i in range(4):
    la = tc.training_step(*batches[i % 2])
    lb = stepd(*batches[i % 2])
    print("losses eager", [float(x) for x in la], "graph", [float(x) for x in lb])
diff = (tc.flat.flat - td.flat.flat).abs().max().item()
print("max |param eager - param graph| after 7 steps:", diff, "adam state", tc.optim.state.tolist(), td.optim.state.tolist())
