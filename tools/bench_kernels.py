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
