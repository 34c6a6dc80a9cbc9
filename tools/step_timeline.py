"""Per-kernel device time of one training step from a CUPTI trace (torch.profiler): the graph-replayed,
stream-overlapped step (MODE=graph, default) or the eager single-stream step (MODE=eager).  Prints, per kernel
name, launches per step, summed duration per step and mean duration; plus the union of busy intervals (device time
with at least one kernel running) and the wall time per step.  Durations of kernels launched with programmatic
dependent launch include the time they wait in griddepcontrol.wait for their predecessor (SININN_PDL=0 removes it)."""
import os, sys, json, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R
from sin_inn_b200 import archs, train
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
P, B = int(os.environ.get("P", 256)), int(os.environ.get("B", 32))
MODE = os.environ.get("MODE", "graph")
STEPS = int(os.environ.get("STEPS", 3))
opt = R.make_opt(scale=4, num_coupling=4, lr_window=10, precision=os.environ.get("PRECISION", "bf16"))
torch.manual_seed(0)
net = archs.UncondSRFlow(3, P, P, opt).to(dev)
tr = train.SingleVideoTrainer(net, opt)
batch = tuple(t.to(dev) for t in R.synthetic_batch(opt, B, P, P, seed=0))
if MODE == "graph":
    step = tr.capture(*batch, warmup=3)
else:
    tr.overlap = False
    net.plan().side_wgrad = False
    step = tr.training_step
    for _ in range(3):
        step(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        step(*batch)
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", f"timeline_{MODE}.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
by = collections.defaultdict(lambda: [0, 0.0])
iv = []
for e in ev:
    name = e["name"].split("(")[0].replace("sininn::", "").replace("void ", "")[:60]
    by[name][0] += 1
    by[name][1] += e["dur"]
    iv.append((e["ts"], e["ts"] + e["dur"]))
iv.sort()
busy, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += (cur_e - cur_s) if cur_e is not None else 0.0
wall = (iv[-1][1] - iv[0][0]) if iv else 0.0
tot = sum(v[1] for v in by.values())
print(f"mode {MODE}: {len(ev) / STEPS:.0f} device activities/step, wall {wall / STEPS / 1e3:.3f} ms/step, "
      f"busy union {busy / STEPS / 1e3:.3f} ms/step, summed durations {tot / STEPS / 1e3:.3f} ms/step")
print("| kernel | launches/step | ms/step | share | mean us |")
print("|---|---:|---:|---:|---:|")
for name, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n / STEPS:.0f} | {us / STEPS / 1e3:.3f} | {100 * us / tot:.1f}% | {us / n:.1f} |")
os.remove(path)
