"""Phase stamps (clock64) of pair 0 of the CTA-pair weight-gradient kernel for the bench shapes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import kernels as K
from sin_inn_b200._lib import load

DEV = "cuda"
B = int(os.environ.get("B", 32))
buf = torch.zeros(4 * 512, dtype=torch.int64, device=DEV)
NAMES = ["entry", "prologue", "dep_wait", "producer_done", "first_stage", "last_mma", "acc_ready", "stored"]
for lvl, hw, c in (("L0", 64, 48), ("L1", 32, 192)):
    for taps in (9, 1):
        for cin, cout in ((c // 2, 256), (256, c)):
            npix = B * hw * hw
            x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
            dy = torch.randn(npix, cout, device=DEV).to(torch.bfloat16)
            k = 3 if taps == 9 else 1
            dw = torch.empty(cout, cin, k, k, device=DEV)
            db = torch.empty(cout, device=DEV)
            for _ in range(2):
                K.wgrad(x, dy, (B, hw, hw), taps, dw, tensor_core=True, dbias=db)
            torch.cuda.synchronize()
            load().sininn_debug_set_trace(buf.data_ptr())
            K.wgrad(x, dy, (B, hw, hw), taps, dw, tensor_core=True, dbias=db)
            torch.cuda.synchronize()
            load().sininn_debug_set_trace(None)
            t = buf[1536:1544].cpu().tolist()
            rel = [v - t[0] for v in t]
            print(f"{lvl} {k}x{k} {cin:3d}->{cout:3d}: " + "  ".join(f"{n}={r}" for n, r in zip(NAMES, rel)), flush=True)
