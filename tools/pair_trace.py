"""In-kernel timeline of the CTA-pair 3x3 conv kernel (CTA 0): clock64 stamps of the producer / MMA / epilogue roles."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import kernels as K
from sin_inn_b200._lib import load

DEV = "cuda"
B = 32
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def run(name, hw, cin, cout, out_dtype, relu=False, bits=False, masked=False):
    npix = B * hw * hw
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device=DEV) * 0.05
    wp = K.pack_weight(w, 0, torch.bfloat16, (cout + 15) // 16 * 16, (cin + 15) // 16 * 16)
    out = torch.zeros(npix, cout, dtype=out_dtype, device=DEV)
    bo = torch.empty(npix, (cout + 31) // 32, dtype=torch.int32, device=DEV) if bits else None
    mb = torch.randint(-2**31, 2**31 - 1, (npix, (cout + 31) // 32), dtype=torch.int32, device=DEV) if masked else None
    bias = None if masked else torch.randn(cout, device=DEV)
    fn = lambda: K.conv(x, wp, (B, hw, hw), cout, out, bias=bias, act=1 if relu else 0, tensor_core=True, bits_out=bo, mask_bits=mb)
    fn(); fn()
    trace = torch.zeros(4, 512, dtype=torch.int64, device=DEV)
    flush.zero_()
    load().sininn_debug_set_trace(trace.data_ptr())
    fn()
    torch.cuda.synchronize()
    load().sininn_debug_set_trace(0)
    t = trace.cpu()
    t0 = int(t[t > 0].min())
    print(f"== {name}: stamps in cycles since the first stamp of CTA 0")
    for role, nm in enumerate(("producer", "mma", "epilogue")):
        v = [int(a) - t0 for a in t[role][:256] if a > 0]
        print(f"  {nm:9s} n={len(v):3d}: " + " ".join(str(a) for a in v[:64]))
    for k in range(3):
        v = [int(a) - t0 for a in t[2][256 + 16 * k:256 + 16 * k + 16] if a > 0]
        print(f"  epilogue tile {k} inner stamps (per slab: start, store-read done, tmem loaded, math+sts done, tma issued): " + " ".join(map(str, v)))


which = sys.argv[1:] or ["c2", "c1", "d2", "l1"]
if "c2" in which: run("L0 conv2 fprop 256->48 (resident)", 64, 256, 48, torch.float32)
if "c1" in which: run("L0 conv1 fprop 24->256 + ReLU + sign bits (resident, n_tile 128)", 64, 24, 256, torch.bfloat16, relu=True, bits=True)
if "d2" in which: run("L0 masked dgrad 48->256 (resident, n_tile 128)", 64, 48, 256, torch.bfloat16, masked=True)
if "l1" in which: run("L1 conv2 fprop 256->192 (streamed)", 32, 256, 192, torch.float32)
