"""Single-kernel driver for ncu captures: python tools/ncu_one.py <case>"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sin_inn_b200 import kernels as K
DEV = "cuda"
case = sys.argv[1] if len(sys.argv) > 1 else "c1x1"
B = 32
shapes = {  # hw, cin, cout, taps, out dtype, relu
    "c1x1": (64, 24, 256, 1, torch.bfloat16, True),
    "c3x3_conv1": (64, 24, 256, 9, torch.bfloat16, True),
    "c3x3_conv2": (64, 256, 48, 9, torch.float32, False),
    "l1_conv2": (32, 256, 192, 9, torch.float32, False),
}
if case.startswith("wg_"):
    hw, cin, cout, taps = {"wg_l0c2": (64, 256, 48, 9), "wg_l1c2": (32, 256, 192, 9), "wg_l0c1": (64, 24, 256, 9)}[case]
    npix = B * hw * hw
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16); dy = torch.randn(npix, cout, device=DEV).to(torch.bfloat16)
    dw = torch.empty(cout, cin, 3, 3, device=DEV)
    for _ in range(3):
        K.wgrad(x, dy, (B, hw, hw), taps, dw, tensor_core=True)
else:
    hw, cin, cout, taps, odt, relu = shapes[case]
    npix = B * hw * hw
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    w = torch.randn(cout, cin, k, k, device=DEV) * 0.05
    wp = K.pack_weight(w, 0, torch.bfloat16, (cout + 15) // 16 * 16, (cin + 15) // 16 * 16)
    bias = torch.randn(cout, device=DEV)
    out = torch.zeros(npix, cout, dtype=odt, device=DEV)
    for _ in range(3):
        K.conv(x, wp, (B, hw, hw), cout, out, bias=bias, act=1 if relu else 0, tensor_core=True)
torch.cuda.synchronize()
print("done", case)
