"""Per-shape timing of the subnet GEMM kernels at the bench workload (B=32, 256x256 patches): CUDA events,
L2 flushed between repetitions.  Prints TFLOP/s per launch shape."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import kernels as K

DEV = "cuda"
B = int(os.environ.get("B", 32))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

def conv_case(name, hw, cin, cout, taps, out_dtype, relu=False, mask=False, acc=False, tc=True):
    npix = B * hw * hw
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    w = torch.randn(cout, cin, k, k, device=DEV) * 0.05
    wp = K.pack_weight(w, 0, torch.bfloat16, (cout + 15) // 16 * 16, (cin + 15) // 16 * 16)
    bias = torch.randn(cout, device=DEV)
    out = torch.zeros(npix, cout, dtype=out_dtype, device=DEV)
    mb = torch.randint(-2**31, 2**31 - 1, (npix, (cout + 31) // 32), dtype=torch.int32, device=DEV) if mask else None
    bo = torch.empty(npix, (cout + 31) // 32, dtype=torch.int32, device=DEV) if relu else None
    fn = lambda: K.conv(x, wp, (B, hw, hw), cout, out, bias=None if (mask or acc) else bias, act=1 if relu else 0,
                        mask_bits=mb, bits_out=bo, accumulate=acc, tensor_core=tc)
    ms = timeit(fn)
    fl = 2.0 * npix * cin * cout * taps
    print(f"{name:34s} M={npix:7d} K={cin*taps:5d} N={cout:4d}  {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s")
    return ms

def wgrad_case(name, hw, cin, cout, taps):
    npix = B * hw * hw
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
    dy = torch.randn(npix, cout, device=DEV).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    dw = torch.empty(cout, cin, k, k, device=DEV)
    ms = timeit(lambda: K.wgrad(x, dy, (B, hw, hw), taps, dw, tensor_core=True))
    fl = 2.0 * npix * cin * cout * taps
    print(f"{name:34s} M={npix:7d} K={cin*taps:5d} N={cout:4d}  {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s")
    return ms

def kernel_times(fn, reps=5):
    """Device duration of every kernel fn() launches (CUPTI through torch.profiler), L2 flushed before each call:
    {kernel name: mean us}."""
    from torch.profiler import profile, ProfilerActivity
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
    out = {}
    for ev in prof.key_averages():
        if "Memset" in ev.key or "fill" in ev.key.lower():
            continue
        out[ev.key.split("(")[0][-40:]] = ev.device_time_total / max(ev.count, 1)
    return out


def wgrad_check(name, hw, cin, cout, taps, bias=True):
    """Same shapes, timed with the fused bias gradient, plus the relative error against torch's fp32 weight gradient
    of the same bf16-rounded operands (cuDNN, TF32 off) and a bit-equality check of two runs."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    npix = B * hw * hw
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
    dy = torch.randn(npix, cout, device=DEV).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    dw = torch.empty(cout, cin, k, k, device=DEV)
    db = torch.empty(cout, device=DEV) if bias else None
    fn = lambda: K.wgrad(x, dy, (B, hw, hw), taps, dw, tensor_core=True, dbias=db)
    ms = timeit(fn)
    first = dw.clone()
    fn(); torch.cuda.synchronize()
    same = torch.equal(first, dw)
    xn = x.float().view(B, hw, hw, cin).permute(0, 3, 1, 2).contiguous()
    dyn = dy.float().view(B, hw, hw, cout).permute(0, 3, 1, 2).contiguous()
    ref = torch.nn.grad.conv2d_weight(xn, (cout, cin, k, k), dyn, padding=k // 2)
    err = float((dw - ref).norm() / ref.norm())
    berr = float((db - dyn.sum((0, 2, 3))).norm() / dyn.sum((0, 2, 3)).norm()) if bias else 0.0
    fl = 2.0 * npix * cin * cout * taps
    kt = kernel_times(fn) if os.environ.get("KT", "1") != "0" else {}
    print(f"{name:24s} {cin:3d}x{cout:3d}x{taps}  {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s  rel_err {err:.2e}  bias_err {berr:.2e}  "
          f"deterministic {same}  " + "  ".join(f"{k}={v:.1f}us" for k, v in kt.items()), flush=True)
    return ms


def fused1x1_case(name, hw, cin, cout, keep):
    npix = B * hw * hw
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
    w1 = torch.randn(256, cin, 1, 1, device=DEV) * 0.05
    w2 = torch.randn(cout, 256, 1, 1, device=DEV) * 0.05
    w1p = K.pack_weight(w1, 0, torch.bfloat16, 256, (cin + 15) // 16 * 16)
    w2p = K.pack_weight(w2, 0, torch.bfloat16, (cout + 15) // 16 * 16, 256)
    b1, b2 = torch.randn(256, device=DEV), torch.randn(cout, device=DEV)
    out = torch.zeros(npix, cout, device=DEV)
    h = torch.empty(npix, 256, dtype=torch.bfloat16, device=DEV) if keep else None
    bits = torch.empty(npix, 8, dtype=torch.int32, device=DEV) if keep else None
    ms = timeit(lambda: K.subnet1x1_fwd(x, w1p, b1, w2p, b2, out, h_out=h, bits_out=bits))
    fl = 2.0 * npix * 256 * (cin + cout)
    print(f"{name:34s} M={npix:7d} {cin:3d}->256->{cout:3d} keep={int(keep)}  {ms*1e3:8.1f} us  {fl/ms/1e9:8.1f} TFLOP/s")
    return ms


if __name__ == "__main__":
    if os.environ.get("ONLY") == "fused1x1":
        for lvl, hw, c in (("L0", 64, 48), ("L1", 32, 192)):
            for keep in (False, True):
                fused1x1_case(f"{lvl} fused 1x1 subnet fwd", hw, c // 2, c, keep)
        sys.exit(0)
    if os.environ.get("ONLY") == "wgrad":
        tot = 0.0
        for lvl, hw, c in (("L0", 64, 48), ("L1", 32, 192)):
            for taps in (9, 1):
                t = f"{lvl} {'3x3' if taps == 9 else '1x1'}"
                tot += wgrad_check(f"{t} conv1 wgrad+bias", hw, c // 2, 256, taps)
                tot += wgrad_check(f"{t} conv2 wgrad+bias", hw, 256, c, taps)
        print(f"sum x8 = {8 * tot:.3f} ms of weight-gradient launches per train step (B={B})")
        sys.exit(0)
    bf, f32 = torch.bfloat16, torch.float32
    tot = 0.0
    for lvl, hw, c in (("L0", 64, 48), ("L1", 32, 192)):
        h = c // 2
        for taps in (9, 1):
            t = f"{lvl} {'3x3' if taps == 9 else '1x1'}"
            a = conv_case(f"{t} conv1 fprop  {h}->256 relu", hw, h, 256, taps, bf, relu=True)
            b_ = conv_case(f"{t} conv2 fprop  256->{c}", hw, 256, c, taps, f32)
            c_ = conv_case(f"{t} conv2 dgrad  {c}->256 mask", hw, c, 256, taps, bf, mask=True)
            d = conv_case(f"{t} conv1 dgrad  256->{h} acc", hw, 256, h, taps, f32, acc=True)
            e = wgrad_case(f"{t} conv1 wgrad", hw, h, 256, taps)
            f = wgrad_case(f"{t} conv2 wgrad", hw, 256, c, taps)
            # per train step: 2 blocks of this kind per level x 2 subnets x 2 directions; fprop runs twice (recompute)
            tot += 8 * (2 * (a + b_) + c_ + d + e + f)
    print(f"sum over one train step (B={B}): {tot:.2f} ms of subnet GEMM kernels")
