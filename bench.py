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
def cpu_reference_run(steps, warmup, batch, threads=None, patch=WORKLOAD["patch"], scale=WORKLOAD["scale"],
                      lr_window=WORKLOAD["lr_window"]):
    """The reference's own CPU implementation of the path: its archs.py restated in oracle/ref_torch.py
    (the reference cannot travel to the GPU box and needs the un-installable FrEIA), fp32, all host threads."""
    from oracle import ref_torch as R
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    opt = R.make_opt(scale=scale, num_coupling=WORKLOAD["num_coupling"], lr_window=lr_window)
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
    from sin_inn_b200 import archs, kernels, train
    from sin_inn_b200 import config as R         # make_opt / synthetic_batch (the oracle is only used by the CPU legs)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P, B = WORKLOAD["patch"], args.batch
    opt = R.make_opt(scale=WORKLOAD["scale"], num_coupling=WORKLOAD["num_coupling"], lr_window=WORKLOAD["lr_window"],
                     precision=args.precision, tensor_core=not args.no_tensor_core, activations=args.activations)
    torch.manual_seed(0)
    net = archs.UncondSRFlow(3, P, P, opt).to(dev)
    trainer = train.SingleVideoTrainer(net, opt, world_size=world)
    trainer.broadcast_params()
    # synthetic data: a pool of pinned host batches (per-rank seed), copied H2D inside the e2e region
    # (z is drawn on the device inside the step, as lit_wrapper.py:41 does: a batch is (hr, lr))
    pool = []
    for i in range(2):
        hr, lr, _ = R.synthetic_batch(opt, B, P, P, seed=1000 * rank + i, with_z=False)
        pool.append(tuple(t.pin_memory() for t in (hr, lr)))
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
            graphed = trainer.capture(*dev_batches[0], None, warmup=2)
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
        # Eager launching is host-bound (~10 ms of enqueue work per step): without a head start the GPU idles between
        # launches and every event pair would also time the host's gap.  A 40 ms spin kernel goes first, the host queues the
        # step behind it, and the GPU then runs the launches back to back: the pairs time kernel + launch latency only.
        torch.cuda._sleep(int(0.04 * 1.9e9))
        trainer.training_step(*dev_batches[i % 2])     # eager: events cannot be recorded inside a graph replay
        torch.cuda.synchronize()
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
        inf = inference_1080p(net, opt, dev, args.infer_batch, iters=args.infer_iters, use_graph=not args.no_graph, world=world, rank=rank)
    t = torch.tensor([ms, ms_e2e] + ([inf["fwd_ms"], inf["inv_ms"], inf["e2e_ms"]] if inf else [0.0, 0.0, 0.0]), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, inf_fwd_ms, inf_inv_ms, inf_e2e_ms = t.tolist()
    # what the differentiable passes kept for their backward (engine.EngineConfig.activations, resolved by the plan)
    plan_, cfg_ = net.plan(), net.engine_config
    stores = any(plan_._store_choice.values()) if cfg_.activations == "auto" else cfg_.activations == "store"
    f_ = 2 * opt.scale
    stash_b = (plan_.stash_bytes((B, 3, P, P), False, cfg_)
               + plan_.stash_bytes((B, opt.lr_dims + opt.z_dims, P // f_, P // f_), True, cfg_)) if stores else 0
    extras = {}
    if world == 1 and not args.no_extras:
        graphed = feeder = None
        trainer._graph = None
        del trainer, net
        torch.cuda.empty_cache()
        extras = extra_configs(dev, args)
    if rank == 0:
        pk = peaks()
        value = world * B * args.steps / (ms / 1e3)
        e2e = world * B * args.steps / (ms_e2e / 1e3)
        zero = {"ms": 0.0, "flops": 0.0, "n": 0, "bytes": 0.0}
        # Tensor-bound kernel families of the eager single-stream pass: algorithmic FLOPs as passed by the wrappers
        # (2 * pixels * Cin * Cout * taps per launch; the recompute launches inside the backward pass are executed work
        # and are counted as such) / summed CUDA-event time.  The family with the largest share of the step is the one the
        # roofline object describes; every family is listed in roofline.families.
        tens = {k: prof.get(k, zero) for k in ("conv3x3", "wgrad", "subnet1x1", "subnet1x1_bwd", "conv1x1")}
        tot_ms = sum(v["ms"] for v in prof.values())
        top = max(tens, key=lambda k: tens[k]["ms"])
        names = {"conv3x3": "conv_tc_pair_kernel (3x3 subnet convolutions: fprop, dgrad; their re-evaluation when activations='recompute')",
                 "wgrad": "wgrad_pair_kernel + wgrad_reduce_kernel (weight and bias gradients, grouped per coupling block)",
                 "subnet1x1": "subnet1x1_fwd_kernel (fused 1x1 subnets: forward, data gradients; re-evaluation when activations='recompute')",
                 "subnet1x1_bwd": "subnet1x1_bwd_kernel + wgrad_reduce_kernel (fused backward of the level-0 1x1 subnets: hidden activation "
                                  "re-evaluated on chip, input gradient, both weight / bias gradients; executed FLOPs incl. the re-evaluation)",
                 "conv1x1": "conv_tc_kernel (1x1 convolutions outside the fused kernel)"}

        def fam(v):
            tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["ms"] > 0 else 0.0
            return {"achieved": tf, "frac": tf / pk["tf_burst"], "frac_of_sustained": tf / pk["tf_sustained"],
                    "ms_per_step": v["ms"] / args.steps, "launches_per_step": v["n"] / args.steps,
                    "share_of_step": v["ms"] / tot_ms if tot_ms > 0 else 0.0}
        fams = {k: fam(v) for k, v in tens.items()}
        tc_ms, tc_fl = sum(v["ms"] for v in tens.values()), sum(v["flops"] for v in tens.values())
        traffic, tnote = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic_r2.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            ent = tj.get("families", {}).get(top)
            if ent:
                traffic, tnote = ent["dram_bytes_per_launch"], ent["note"]
        alg_flops_step = 6 * 2 * conv_macs_per_patch(P) * B

        # algorithmic-only: 4 of the 6 conv3x3/1x1 passes per direction are algorithmic (fprop, dgrad; + wgrad), the
        # recomputed fprop is not: algorithmic FLOPs of the step / time of ALL tensor-bound launches
        alg_only = alg_flops_step * args.steps / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0
        line = {
            "metric": "INN train-step patches/sec", "value": value, "unit": "patches/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_spread": spread, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "fp32": "f32", "fp32tc": "f32 (bf16 split operands)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": f"SRF scale4 c4 lr_window10 {P}x{P} train step, batch {B}/GPU", **WORKLOAD,
                       "batch_per_gpu": B, "global_batch": B * world, "precision": args.precision, "tensor_core": not args.no_tensor_core,
                       "parallelism": f"dp{world}", "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2",
                       "backward": "trunk restored block by block from the exact inverse (never stored); coupling-subnet internals "
                                   + ("kept from the value pass" if stores else "re-evaluated during backward"),
                       "activations": args.activations, "subnet_state_bytes_per_step": stash_b,
                       "cuda_graph": used_graph, "z": "drawn on the device inside the step"},
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": fams[top]["achieved"], "peak": pk["tf_burst"], "unit": "TFLOP/s",
                         "frac": fams[top]["frac"], "frac_of_sustained": fams[top]["frac_of_sustained"],
                         "peak_sustained": pk["tf_sustained"], "traffic": traffic, "traffic_note": tnote,
                         "peak_source": pk["source"] + " (burst bf16; the sustained figure is alongside)",
                         "kernel": names[top], "kernel_ms_per_step": fams[top]["ms_per_step"],
                         "kernel_launches": tens[top]["n"], "kernel_share_of_step": fams[top]["share_of_step"],
                         "families": fams,
                         "all_subnet_gemms": {"achieved": tc_fl / (tc_ms / 1e3) / 1e12 if tc_ms > 0 else 0.0,
                                              "frac": (tc_fl / (tc_ms / 1e3) / 1e12 / pk["tf_burst"]) if tc_ms > 0 else 0.0,
                                              "ms_per_step": tc_ms / args.steps},
                         "algorithmic_only_tflops": alg_only, "algorithmic_only_frac": alg_only / pk["tf_burst"],
                         "whole_step_algorithmic_tflops": alg_flops_step / (ms / args.steps / 1e3) / 1e12,
                         "whole_step_frac": alg_flops_step / (ms / args.steps / 1e3) / 1e12 / pk["tf_burst"],
                         "timing_note": "family times come from a second, eager single-stream pass of the same K steps with a CUDA-event "
                                        "pair around every launch (launch gaps included); the headline step is the graph-replayed, "
                                        "stream-overlapped schedule of the same launches, tested bit-identical"},
            # bandwidth-bound kernel families: algorithmic bytes of their launches / CUDA-event time, vs measured HBM peak
            "roofline_hbm": {fam_: {"achieved": prof[fam_]["bytes"] / (prof[fam_]["ms"] / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                    "frac": prof[fam_]["bytes"] / (prof[fam_]["ms"] / 1e3) / 1e9 / pk["hbm"],
                                    "launches": prof[fam_]["n"], "ms_per_step": prof[fam_]["ms"] / args.steps}
                             for fam_ in ("coupling_bwd", "coupling", "permute", "resample", "layout", "split")
                             if fam_ in prof and prof[fam_]["ms"] > 0 and prof[fam_]["bytes"] > 0},
            "clocks": sampler.summary(),
            "profile_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
        }
        if inf:
            fr = world * inf["frames"]
            # bf16 round trip inverse(forward(x)): the inverse re-evaluates every subnet on bf16-rounded inputs that differ in
            # the last fp32 bits, so single bf16 roundings flip; bounds asserted here: max 6e-2, mean 1e-3
            rt_ok = inf["roundtrip"] <= 6e-2 and inf["roundtrip_mean"] <= 1e-3
            line["inference_1080p"] = {
                "workload": "SRF scale4 c4 1920x1080 frames, forward + inverse, no_grad, micro-batch %d/GPU (two in flight), frame-sharded" % args.infer_batch,
                "fwd_inv_frames_per_s": fr / ((inf_fwd_ms + inf_inv_ms) / 1e3),
                "fwd_frames_per_s": fr / (inf_fwd_ms / 1e3), "inv_frames_per_s": fr / (inf_inv_ms / 1e3),
                "frames_timed_per_gpu": inf["frames"], "roundtrip_max_abs_err": inf["roundtrip"],
                "roundtrip_mean_abs_err": inf["roundtrip_mean"], "roundtrip_within_bounds": rt_ok,
                "algorithmic_tflops": 2 * 382.2e9 * fr / ((inf_fwd_ms + inf_inv_ms) / 1e3) / 1e12,
                # BASELINE.json configs[3] end to end: a 120-frame clip sharded over the ranks; per frame the LR window comes
                # from pinned host memory, z is drawn on the device (lit_wrapper.py:110), the inverse runs as a replayed graph,
                # the frame is quantised to uint8 HWC on the device and copied to pinned host memory (lit_wrapper.py:117-121)
                "e2e_120_frames": {"frames_per_s": world * inf["e2e_frames"] / (inf_e2e_ms / 1e3), "frames_per_gpu": inf["e2e_frames"],
                                   "h2d_bytes_per_frame": inf["e2e_h2d"], "d2h_bytes_per_frame": inf["e2e_d2h"],
                                   "direction": "inverse (LR, z) -> HR = the reference's infer"}}
            assert rt_ok, f"bf16 round trip out of bounds: max {inf['roundtrip']}, mean {inf['roundtrip_mean']}"
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=3, warmup=1, batch=2)
            line["cpu_baseline"] = {"value": r["value"], "unit": "patches/s", "cores": r["cores"], "kind": "port",
                                    "sample": "3 steps x 2 patches of 256x256, same model, oracle port of archs.py on host cores"}
            r0 = cpu_reference_run(steps=5, warmup=1, batch=4, patch=64, scale=2, lr_window=1)
            line["config0_cpu"] = {"workload": "BASELINE.json configs[0]: SRF scale 2, 4 couplings, 8-frame 64x64 clip (batch 4), fp32, "
                                               "oracle port on host cores, full step", "patches_per_s": r0["value"],
                                   "ms_per_step": r0["ms_per_step"], "cores": r0["cores"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured step holds graph-captured NCCL work: release it (as the inference leg does) and let every rank get here
        # before the communicator is torn down -- destroying the communicator under a live graph can block (a multi-rank
        # run with --no-inference hung at exit in round 2 and ran into its time limit).
        graphed = feeder = None
        if "trainer" in locals() and trainer is not None:
            trainer._graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        dist.destroy_process_group()


def _time_steps(step, n, warm=2):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def extra_configs(dev, args):
    """Numbers for the BASELINE.json configurations the headline does not cover (N = 1 only; a few seconds each):
    configs[1] "fp32 vs bf16" (both fp32 paths), configs[4] deep variant (throughput + peak memory), the IRN architecture."""
    from sin_inn_b200 import archs, train
    from sin_inn_b200 import config as R
    out = {}
    P, B = WORKLOAD["patch"], args.batch

    def trainer_for(opt, ctor, patch, batch, graph=True):
        torch.manual_seed(0)
        net = ctor(3, patch, patch, opt).to(dev)
        if opt.architecture == "IRN":             # conv5 is zero-initialised in the reference: a fresh IRN is the identity
            g = torch.Generator(device="cpu").manual_seed(1)
            for m in net.modules():
                if isinstance(m, archs.DenseBlock):
                    m.conv5.weight.data.copy_(0.02 * torch.randn(m.conv5.weight.shape, generator=g))
        tr = train.SingleVideoTrainer(net, opt)
        hr, lr, _ = (t.to(dev) if t is not None else None for t in R.synthetic_batch(opt, batch, patch, patch, seed=0, with_z=False))
        step = None
        if graph and not args.no_graph:
            try:
                g_ = tr.capture(hr, lr, None, warmup=2)
                step = lambda: g_(hr, lr)
            except Exception as e:
                print(f"[bench] capture failed for an extra config ({e}); eager", file=sys.stderr, flush=True)
        if step is None:
            step = lambda: tr.training_step(hr, lr)
        return tr, step

    for prec, key, batch, steps in (("fp32tc", "fp32_path", B, 5), ("fp32", "fp32_cuda_core_path", 8, 2)):
        opt = R.make_opt(scale=WORKLOAD["scale"], num_coupling=WORKLOAD["num_coupling"], lr_window=WORKLOAD["lr_window"], precision=prec)
        tr, step = trainer_for(opt, archs.UncondSRFlow, P, batch)
        msv = _time_steps(step, steps)
        out[key] = {"workload": f"configs[1] in precision {prec!r}: " + ("fp32 activations, subnet GEMMs on the tensor cores over bf16 hi/mid/lo "
                    "split operands" if prec == "fp32tc" else "CUDA-core fp32 kernels, the reference-accurate path"),
                    "batch": batch, "value": batch / (msv / 1e3), "unit": "patches/s", "ms_per_step": msv}
        del tr, step
        torch.cuda.empty_cache()
    # configs[1] with the subnets re-evaluated during backward instead of kept (activations="recompute")
    def mem_and_time(opt, patch, batch, steps, graph):
        tr, step = trainer_for(opt, archs.UncondSRFlow, patch, batch, graph=graph)
        step()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        msv = _time_steps(step, steps, warm=1)
        peak = torch.cuda.max_memory_allocated() - base
        del tr, step
        torch.cuda.empty_cache()
        return msv, peak
    opt = R.make_opt(scale=WORKLOAD["scale"], num_coupling=WORKLOAD["num_coupling"], lr_window=WORKLOAD["lr_window"], precision="bf16",
                     activations="recompute")
    msv, peak = mem_and_time(opt, P, B, 5, False)
    out["recompute_variant"] = {"workload": "configs[1], bf16, eager step, activations='recompute': every coupling subnet is re-evaluated "
                                            "from the restored trunk during backward; nothing but the network output is kept",
                                "value": B / (msv / 1e3), "unit": "patches/s", "ms_per_step": msv, "peak_extra_memory_GiB": peak / 2 ** 30}
    # configs[4]: deep variant, 512x512 patches: throughput and peak memory in both modes
    out["deep_variant"] = {"workload": "configs[4]: SRF scale 4, 8 couplings per level, hidden 512, 512x512 patches, batch 8, bf16, eager step"}
    for mode in ("recompute", "store"):
        opt = R.make_opt(scale=4, num_coupling=8, lr_window=10, precision="bf16", hidden=512, activations=mode)
        msv, peak = mem_and_time(opt, 512, 8, 3, False)
        out["deep_variant"][mode] = {"value": 8 / (msv / 1e3), "unit": "patches/s", "ms_per_step": msv, "peak_extra_memory_GiB": peak / 2 ** 30}
    out["deep_variant"]["value"] = out["deep_variant"]["store"]["value"]
    out["deep_variant"]["unit"] = "patches/s"
    out["deep_variant"]["note"] = ("'recompute' = fused recompute-from-inverse backward (only the network output is kept: the memory figure is "
                                   "the working set of one coupling block); 'store' keeps the subnets' operand copy, hidden activation, sign "
                                   "bits and output per block (the trunk is still rebuilt from the inverse)")
    opt = R.make_opt(scale=WORKLOAD["scale"], num_coupling=WORKLOAD["num_coupling"], lr_window=WORKLOAD["lr_window"], precision="bf16",
                     architecture="IRN")
    tr, step = trainer_for(opt, archs.InvRescaleNet, P, B)
    msv = _time_steps(step, 5)
    out["irn_arch"] = {"workload": f"InvRescaleNet (archs.py:201-233) scale 4, 4 InvBlockExp per level, {P}x{P}, batch {B}, bf16",
                       "value": B / (msv / 1e3), "unit": "patches/s", "ms_per_step": msv,
                       "algorithmic_tflops": 6 * 20.18e9 * B / (msv / 1e3) / 1e12}
    del tr, step
    torch.cuda.empty_cache()
    return out


def inference_1080p(net, opt, dev, micro_batch, iters, use_graph=True, world=1, rank=0):
    """Times `iters` micro-batches of 1920x1080 frames through net(x) and net(lr_z, rev=True) (CUDA events); each
    direction is one replayed CUDA graph (train.GraphedInference) unless use_graph is False.  Then BASELINE.json
    configs[3] end to end: this rank's share of a 120-frame clip, host LR windows in, uint8 frames out."""
    from sin_inn_b200 import engine, train
    H, W = 1080, 1920
    g = torch.Generator(device="cpu").manual_seed(7)
    hr = torch.rand(micro_batch, 3, H, W, generator=g).to(dev)
    out = {}
    with torch.no_grad():
        lrz = net(hr)
        back = net(lrz, rev=True)
        out["roundtrip"] = float((back - hr).abs().max())
        out["roundtrip_mean"] = float((back - hr).abs().mean())
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
        # ---- configs[3] end to end: LR windows of this rank's frames from pinned host memory -> device, z drawn on the
        # device (temp 0.8), inverse pass (graph replay over static lr / z-free LatentInput), uint8 HWC on the device,
        # one pinned D2H copy per micro-batch; H2D of micro-batch i+1 overlaps the compute of i (side stream)
        del gfs, gis
        torch.cuda.empty_cache()
        frames = list(train.shard_frames(120, rank, world))
        mb = micro_batch
        n_mb = (len(frames) + mb - 1) // mb
        h, w = H // (2 * opt.scale), W // (2 * opt.scale)
        host_lr = [torch.rand(mb, opt.lr_dims, h, w, generator=g).pin_memory() for _ in range(2)]
        host_out = [torch.empty(mb, H, W, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
        slots = [torch.empty(mb, opt.lr_dims, h, w, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        step_state = torch.zeros(3, dtype=torch.int32, device=dev)          # advances the z stream per micro-batch

        def run_mb(slot):
            lat = engine.LatentInput(slots[slot], None, z_dims=opt.z_dims, temp=opt.temp, seed=1234 + rank, step_state=step_state)
            return net(lat, rev=True)

        graphs = None
        if use_graph:
            graphs = []
            for slot in range(2):
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(2):
                        run_mb(slot)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph):
                    o = run_mb(slot)
                    q = train.K.quantize_u8_hwc(o)
                graphs.append((gph, q))

        def submit(i):
            with torch.cuda.stream(copy_stream):
                slots[i % 2].copy_(host_lr[i % 2], non_blocking=True)
                ready[i % 2].record(copy_stream)

        def clip():
            submit(0)
            for i in range(n_mb):
                if i + 1 < n_mb:
                    submit(i + 1)
                torch.cuda.current_stream().wait_event(ready[i % 2])
                if graphs is not None:
                    graphs[i % 2][0].replay()
                    q = graphs[i % 2][1]
                else:
                    q = train.K.quantize_u8_hwc(run_mb(i % 2))
                step_state[0:1].add_(1)
                host_out[i % 2].copy_(q, non_blocking=True)
                copy_stream.wait_stream(torch.cuda.current_stream())        # slot i % 2 is reused by micro-batch i + 2
            torch.cuda.current_stream().synchronize()

        clip()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        clip()
        e1.record()
        torch.cuda.synchronize()
        out["e2e_ms"] = e0.elapsed_time(e1)
        out["e2e_frames"] = n_mb * mb
        out["e2e_h2d"] = opt.lr_dims * h * w * 4
        out["e2e_d2h"] = H * W * 3
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp32tc"])
    ap.add_argument("--activations", default="auto", choices=["auto", "store", "recompute"],
                    help="coupling-subnet internals for backward: kept from the value pass, or re-evaluated (engine.EngineConfig)")
    ap.add_argument("--batch", type=int, default=WORKLOAD["batch_per_gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the 1080p forward+inverse measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32 / deep-variant / IRN measurements (N = 1 only)")
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
