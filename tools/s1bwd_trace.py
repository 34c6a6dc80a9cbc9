"""In-kernel timeline of the fused 1x1 subnet backward kernel (subnet1x1_bwd.cu) at the level-0 shape: clock64 stamps of
CTA 0's MMA issuer and of epilogue warp 2 for its first tiles (sininn_debug_set_trace)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import kernels as K
from sin_inn_b200._lib import load
DEV = "cuda"
B = 32
npix, cin, hid, cout = B * 64 * 64, 24, 256, 48
bf = torch.bfloat16
x = torch.randn(npix, cin, device=DEV).to(bf)
da = torch.randn(npix, cout, device=DEV).to(bf)
w1 = torch.randn(hid, cin, 1, 1, device=DEV) * 0.05
w2 = torch.randn(cout, hid, 1, 1, device=DEV) * 0.05
b1 = torch.randn(hid, device=DEV)
dsrc = torch.zeros(npix, cin, device=DEV)
g = [torch.zeros(hid, cin, 1, 1, device=DEV), torch.zeros(hid, device=DEV), torch.zeros(cout, hid, 1, 1, device=DEV), torch.zeros(cout, device=DEV)]
packs = (K.pack_weight(w1, 0, bf, 256, 32), K.pack_weight(w2, 1, bf, 256, 48), K.pack_weight(w1, 1, bf, 32, 256))
run = lambda: K.subnet1x1_bwd(x, da, packs[0], b1, packs[1], packs[2], dsrc, (g[0], True, g[1], True), (g[2], True, g[3], True))
buf = torch.zeros(4 * 512, dtype=torch.int64, device=DEV)
for _ in range(2):
    run()
torch.cuda.synchronize()
load().sininn_debug_set_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
load().sininn_debug_set_trace(None)
t = buf[1600:1856].cpu().tolist()
mma, epi = t[:128], t[128:]
t0 = min(v for v in t if v > 0)
MN = ["loop top", "hacc_free+da", "h_full", "G4 issued", "G1n issued", "dh_full", "G3/G5 issued"]
EN = ["loop top", "d1_full", "h computed", "s_full", "d2_full", "g4_done", "dh arrived", "Hpre loaded"]
for i in range(8):
    m = [v - t0 for v in mma[8 * i:8 * i + 7]]
    e = [v - t0 for v in epi[8 * i:8 * i + 8]]
    if mma[8 * i] == 0:
        break
    print(f"tile {i}  MMA: " + "  ".join(f"{n}={v}" for n, v in zip(MN, m)))
    print(f"        EPI: " + "  ".join(f"{n}={v}" for n, v in zip(EN, e)))
k = [v - t0 for v in epi[120:127]]
print("kernel phases (thread 0 / warp 2): entry, prologue, setup sync, loop end, partials stored, stores drained, exit:", k)
