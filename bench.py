#!/usr/bin/env python
"""Headline benchmark: INN train-step patches/sec (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N --steps K --warmup W]          one process per GPU (torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's CPU path (oracle port) on host cores

Workload = BASELINE.json configs[1]/[2]: UncondSRFlow scale 4, 4 couplings, lr_window 10 (84 + 108 latent
channels), 256x256 patches, batch 32 PER GPU (weak scaling), synthetic U[0,1) frames and N(0,1) latents,
random-init weights (seed 0).  One step = zero_grad + forward + L2 + backward + inverse + L2 + backward
+ gradient all-reduce (N > 1) + Adam (lit_wrapper.py:36-56,76; loss.mmd excluded, lambda 0).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(arch="SRF", scale=4, num_coupling=4, lr_window=10, patch=256, batch_per_gpu=32)


def conv_macs_per_patch(P, scale=4, num_coupling=4, hidden=256):
    """Algorithmic MACs of the coupling subnets for one PxP patch, one direction:
    per GLOW block at C channels both subnets together do hidden*C*k^2*(1/2+1)... = 3/2*... -> 768*C*k^2 at hidden 256."""
    total = 0
    C, px = 12, (P // 2) ** 2
    for _ in range((scale - 1).bit_length()):
        C, px = C * 4, px // 4
        for k in range(num_coupling):
            kk = 9 if k % 2 == 0 else 1
            total += px * 3 * hidden * C * kk
    return total


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 6:
                reasons |= {n for n, v in zip(names, r[2:6]) if v.lower().startswith("active")}
        mx = max((int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference / CPU arm
def cpu_reference_run(steps, warmup, batch, threads=None, patch=WORKLOAD["patch"]):
    """The reference's own CPU implementation of the path: its archs.py restated in oracle/ref_torch.py
    (the reference cannot travel to the GPU box and needs the un-installable FrEIA), fp32, all host threads."""
    from oracle import ref_torch as R
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    opt = R.make_opt(scale=WORKLOAD["scale"], num_coupling=WORKLOAD["num_coupling"], lr_window=WORKLOAD["lr_window"])
    torch.manual_seed(0)
    net = R.build_srf(3, patch, patch, opt)
    optim = R.make_optimizer(net, opt)
    hr, lr, z = R.synthetic_batch(opt, batch, patch, patch, seed=0)
    for _ in range(warmup):
        R.train_step(net, optim, hr, lr, z, opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        R.train_step(net, optim, hr, lr, z, opt)
    dt = time.perf_counter() - t0
    return dict(value=batch * steps / dt, ms_per_step=1e3 * dt / steps, cores=threads, batch=batch)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2
    r = cpu_reference_run(args.steps, args.warmup, batch)
    sample = f"{args.steps} steps x {batch} patches of 256x256 (same model/config as the GPU arm, bounded batch)"
    line = {
        "impl": "reference", "metric": "INN train-step patches/sec", "value": r["value"], "unit": "patches/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SRF scale4 c4 lr_window10 256x256 train step, CPU (oracle port of archs.py, all host threads), "
                               "batch 2 per step", **{**WORKLOAD, "batch_per_gpu": None, "batch_per_step_cpu": batch}},
        "cpu_baseline": {"value": r["value"], "unit": "patches/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    from oracle import ref_torch as R            # only for make_opt/synthetic_batch helpers + cpu_baseline leg
    from sin_inn_b200 import archs, kernels, train

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P, B = WORKLOAD["patch"], args.batch
    opt = R.make_opt(scale=WORKLOAD["scale"], num_coupling=WORKLOAD["num_coupling"], lr_window=WORKLOAD["lr_window"],
                     precision=args.precision, tensor_core=not args.no_tensor_core)
    torch.manual_seed(0)
    net = archs.UncondSRFlow(3, P, P, opt).to(dev)
    trainer = train.SingleVideoTrainer(net, opt, world_size=world)
    trainer.broadcast_params()
    # synthetic data: a pool of pinned host batches (per-rank seed), copied H2D inside the e2e region
    pool = []
    for i in range(2):
        hr, lr, z = R.synthetic_batch(opt, B, P, P, seed=1000 * rank + i)
        pool.append(tuple(t.pin_memory() for t in (hr, lr, z)))
    dev_batches = [tuple(t.to(dev, non_blocking=True) for t in b) for b in pool]
    h2d = sum(t.numel() * 4 for t in pool[0])
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The step is captured once into a CUDA graph (train.SingleVideoTrainer.capture): ~560 kernel launches, the
    # NCCL all-reduce and Adam replay as ONE graph launch.  --no-graph runs the same step eagerly.
    kernels.LAUNCHES = 0
    trainer.training_step(*dev_batches[0])
    launches_per_step = kernels.LAUNCHES
    graphed = None
    if not args.no_graph:
        try:
            graphed = trainer.capture(*dev_batches[0], warmup=2)
        except Exception as e:                      # keep measuring (eagerly) if stream capture is refused on this box
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running the step eagerly", file=sys.stderr, flush=True)
            torch.cuda.synchronize()
            graphed = None

    def step_resident(i):
        if graphed is not None:
            return graphed(*dev_batches[i % 2])       # device-to-device copy into the graph's static inputs + replay
        return trainer.training_step(*dev_batches[i % 2])

    # end-to-end: every step's inputs come from pinned host memory; the copy of batch i+1 is issued on a side
    # stream before step i runs (train.HostBatchFeeder), so H2D overlaps compute; the losses are read back every step
    feeder = train.HostBatchFeeder(pool[0], dev)
    host_losses = torch.zeros(2, dtype=torch.float32).pin_memory()

    def step_e2e(i):
        feeder.submit(pool[(i + 1) % 2], (i + 1) % 2)             # prefetch the next step's batch
        batch = feeder.take(i % 2)
        lf, lb = (graphed or trainer.training_step)(*batch)
        feeder.release(i % 2)
        # D2H read of the step's result: both losses, 8 bytes, one copy + one synchronisation
        pair = graphed.loss_pair if graphed is not None else torch.stack((lf, lb))
        host_losses.copy_(pair, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(host_losses[0] + host_losses[1])

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]     # per-step spread (p10/p50/p90)
    ev0.record()
    marks[0].record()
    for i in range(args.steps):
        step_resident(i)
        marks[i + 1].record()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))
    spread = {"p10": per_step[int(0.1 * (len(per_step) - 1))], "p50": per_step[len(per_step) // 2],
              "p90": per_step[int(0.9 * (len(per_step) - 1) + 0.5)]}
    launches = launches_per_step * args.steps      # kernels executed in the timed region (graph replays included)
    # same K steps again with a CUDA-event pair around every kernel launch (per-family durations for the roofline);
    # kept out of the headline region because ~1300 extra event records per step perturb a launch-dense step
    # (single stream for this pass: with the step's halves overlapped on several streams an event pair would also
    #  time whatever the other streams ran in between)
    saved = (trainer.overlap, net.plan().side_wgrad)
    trainer.overlap, net.plan().side_wgrad = False, False
    trainer.training_step(*dev_batches[0])
    kernels.profile_begin()
    for i in range(args.steps):
        trainer.training_step(*dev_batches[i % 2])     # eager: events cannot be recorded inside a graph replay
    prof = kernels.profile_end()
    trainer.overlap, net.plan().side_wgrad = saved
    sampler.stop_flag = True
    # end-to-end: pinned host inputs -> device, step, loss back to host, every step
    feeder.submit(pool[0], 0)
    for i in range(2):
        step_e2e(i)
    barrier()
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):                      # K steps, K host->device batch copies, K loss read-backs
        step_e2e(i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    # second half of BASELINE.json's metric: full-video inference, 1080p frames through forward (HR -> LR,z) and
    # inverse (LR,z -> HR = the reference's infer, lit_wrapper.py:105-115), frames sharded over the ranks, no
    # communication (configs[3]).  Each rank times its own micro-batches; frames/s = all ranks' frames / max time.
    inf = None
    used_graph = graphed is not None
    if not args.no_inference:
        used_graph = graphed is not None
        graphed = feeder = None
        trainer._graph = None
        trainer.optim.zero_grad()
        torch.cuda.empty_cache()
        inf = inference_1080p(net, opt, dev, args.infer_batch, iters=args.infer_iters, use_graph=not args.no_graph)
    t = torch.tensor([ms, ms_e2e] + ([inf["fwd_ms"], inf["inv_ms"]] if inf else [0.0, 0.0]), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, inf_fwd_ms, inf_inv_ms = t.tolist()
    if rank == 0:
        pk = peaks()
        value = world * B * args.steps / (ms / 1e3)
        e2e = world * B * args.steps / (ms_e2e / 1e3)
        # roofline of the dominant kernel (conv_tc_pair_kernel = every 3x3 subnet convolution, ~30 % of the step's
        # device time): algorithmic FLOPs of its launches / their summed CUDA-event duration in the eager pass
        zero = {"ms": 0.0, "flops": 0.0, "n": 0}
        c3 = prof.get("conv3x3", zero)
        fams = [prof.get(k, zero) for k in ("conv3x3", "conv1x1", "subnet1x1", "wgrad")]
        tc_ms, tc_fl, tc_n = sum(f["ms"] for f in fams), sum(f["flops"] for f in fams), sum(f["n"] for f in fams)
        achieved = c3["flops"] / (c3["ms"] / 1e3) / 1e12 if c3["ms"] > 0 else 0.0
        achieved_all = tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
        traffic, traffic_detail = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            per = [v["dram_read"] + v["dram_write"] for v in tj["shapes"].values()]
            traffic, traffic_detail = sum(per) / len(per), tj
        alg_flops_step = 6 * 2 * conv_macs_per_patch(P) * B
        line = {
            "metric": "INN train-step patches/sec", "value": value, "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_spread": spread, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": f"SRF scale4 c4 lr_window10 {P}x{P} train step, batch {B}/GPU", **WORKLOAD,
                       "batch_per_gpu": B, "global_batch": B * world, "precision": args.precision, "tensor_core": not args.no_tensor_core,
                       "parallelism": f"dp{world}", "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2",
                       "backward": "recompute-from-inverse", "cuda_graph": used_graph},
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tf_sustained"], "traffic": traffic, "peak_source": pk["source"] + " (sustained)",
                         "kernel": "conv_tc_pair_kernel (3x3 subnet convolutions: fprop, recompute, dgrad)",
                         "kernel_ms_per_step": c3["ms"] / args.steps, "kernel_launches": c3["n"],
                         "traffic_note": "mean DRAM bytes per launch over the three ncu --set full captures in profiles/traffic_r1.json",
                         "all_subnet_gemms": {"achieved": achieved_all, "frac": achieved_all / pk["tf_sustained"],
                                              "ms_per_step": tc_ms / args.steps, "launches": tc_n},
                         "whole_step_algorithmic_tflops": alg_flops_step / (ms / args.steps / 1e3) / 1e12},
            # bandwidth-bound kernel families: algorithmic bytes of their launches / CUDA-event time, vs measured HBM peak
            "roofline_hbm": {fam: {"achieved": prof[fam]["bytes"] / (prof[fam]["ms"] / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                   "frac": prof[fam]["bytes"] / (prof[fam]["ms"] / 1e3) / 1e9 / pk["hbm"],
                                   "launches": prof[fam]["n"], "ms_per_step": prof[fam]["ms"] / args.steps}
                             for fam in ("coupling_bwd", "coupling", "permute", "resample", "layout")
                             if fam in prof and prof[fam]["ms"] > 0 and prof[fam]["bytes"] > 0},
            "clocks": sampler.summary(),
            "profile_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        }
        if inf:
            fr = world * inf["frames"]
            line["inference_1080p"] = {
                "workload": "SRF scale4 c4 1920x1080 frames, forward + inverse, no_grad, micro-batch %d/GPU (two in flight), frame-sharded" % args.infer_batch,
                "fwd_inv_frames_per_s": fr / ((inf_fwd_ms + inf_inv_ms) / 1e3),
                "fwd_frames_per_s": fr / (inf_fwd_ms / 1e3), "inv_frames_per_s": fr / (inf_inv_ms / 1e3),
                "frames_timed_per_gpu": inf["frames"], "roundtrip_max_abs_err": inf["roundtrip"],
                "algorithmic_tflops": 2 * 382.2e9 * fr / ((inf_fwd_ms + inf_inv_ms) / 1e3) / 1e12}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=3, warmup=1, batch=2)
            line["cpu_baseline"] = {"value": r["value"], "unit": "patches/s", "cores": r["cores"], "kind": "port",
                                    "sample": "3 steps x 2 patches of 256x256, same model, oracle port of archs.py on host cores"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def inference_1080p(net, opt, dev, micro_batch, iters, use_graph=True):
    """Times `iters` micro-batches of 1920x1080 frames through net(x) and net(lr_z, rev=True) (CUDA events); each
    direction is one replayed CUDA graph (train.GraphedInference) unless use_graph is False."""
    from sin_inn_b200 import train
    H, W = 1080, 1920
    g = torch.Generator(device="cpu").manual_seed(7)
    hr = torch.rand(micro_batch, 3, H, W, generator=g).to(dev)
    out = {}
    with torch.no_grad():
        lrz = net(hr)
        back = net(lrz, rev=True)
        out["roundtrip"] = float((back - hr).abs().max())
        del back
        if use_graph:
            # two independent micro-batches in flight on two streams: the replayed graphs fill each other's kernel tails
            streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
            gfs, gis = [], []
            for st in streams:
                with torch.cuda.stream(st):
                    gfs.append(train.GraphedInference(net, hr, False))
                    gis.append(train.GraphedInference(net, lrz, True))
            torch.cuda.synchronize()

            def both(graphs, x):
                def run():
                    for st, g in zip(streams, graphs):
                        st.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(st):
                            g(x)
                    for st in streams:
                        torch.cuda.current_stream().wait_stream(st)
                return run
            runs = (("fwd_ms", both(gfs, hr)), ("inv_ms", both(gis, lrz)))
            per_call = 2 * micro_batch
        else:
            runs = (("fwd_ms", lambda: net(hr)), ("inv_ms", lambda: net(lrz, rev=True)))
            per_call = micro_batch
        for tag, fn in runs:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out[tag] = e0.elapsed_time(e1)
    out["frames"] = per_call * iters
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32tc"])
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch_per_gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the 1080p forward+inverse measurement")
    ap.add_argument("--infer-batch", type=int, default=2, help="1080p frames per micro-batch and GPU")
    ap.add_argument("--infer-iters", type=int, default=8)
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-tensor-core", action="store_true", help="route the subnet GEMMs to the CUDA-core kernels")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
