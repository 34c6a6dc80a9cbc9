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
if case.startswith("s1x1"):
    # fused 1x1 subnet at level 0: forward (s1x1_fwd), forward storing h + sign bits (s1x1_keep), data gradient (s1x1_grad)
    npix, cin, hid, cout = B * 64 * 64, 24, 256, 48
    x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
    w1 = torch.randn(hid, cin, 1, 1, device=DEV) * 0.05
    w2 = torch.randn(cout, hid, 1, 1, device=DEV) * 0.05
    b1, b2 = torch.randn(hid, device=DEV), torch.randn(cout, device=DEV)
    out = torch.zeros(npix, cout, device=DEV)
    h = torch.empty(npix, hid, dtype=torch.bfloat16, device=DEV)
    bits = torch.zeros(npix, hid // 32, dtype=torch.int32, device=DEV)
    for _ in range(3):
        if case == "s1x1_grad":
            da = torch.randn(npix, cout, device=DEV).to(torch.bfloat16)
            dsrc = torch.zeros(npix, cin, device=DEV)
            K.subnet1x1_fwd(da, K.pack_weight(w2, 1, torch.bfloat16, 256, 48), None, K.pack_weight(w1, 1, torch.bfloat16, 32, 256), None,
                            dsrc, h_out=h, mask_bits=bits, accumulate=True)
        else:
            keep = case == "s1x1_keep"
            K.subnet1x1_fwd(x, K.pack_weight(w1, 0, torch.bfloat16, 256, 32), b1, K.pack_weight(w2, 0, torch.bfloat16, 48, 256), b2, out,
                            h_out=h if keep else None, bits_out=bits if keep else None)
elif case == "bwd1x1":
    # fused 1x1 subnet backward at level 0 (subnet1x1_bwd.cu): recompute of h, dh, input gradient, both weight gradients
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
    for _ in range(3):
        K.subnet1x1_bwd(x, da, packs[0], b1, packs[1], packs[2], dsrc, (g[0], True, g[1], True), (g[2], True, g[3], True))
elif case in ("coupling_bwd", "coupling_apply", "permute", "resample", "colsum"):
    # HBM-bound kernels at the bench workload's level-0 shapes (B=32: 131072 pixels x 48 channels fp32 trunk)
    npix, C, L = B * 64 * 64, 48, 24
    U = torch.randn(npix, C, device=DEV); dU = torch.randn(npix, C, device=DEV)
    a = torch.randn(npix, 2 * L, device=DEV)
    da = torch.empty(npix, 2 * L, dtype=torch.bfloat16, device=DEV)
    cmap = torch.randperm(C, device=DEV).to(torch.int32)
    x4 = torch.randn(B, 12, 128, 128, device=DEV)
    for _ in range(3):
        if case == "coupling_bwd":
            K.coupling_bwd(U[:, :L], dU[:, :L], a[:, :L], a[:, L:], 0, 1.2, 0, da[:, :L], da[:, L:], True)
        elif case == "coupling_apply":
            K.coupling_apply(U[:, :L], a[:, :L], a[:, L:], 0, 1.2, 0, True)
        elif case == "permute":
            K.permute_nhwc(U.view(B, 64, 64, C), cmap, (0, 24))
        elif case == "resample":
            K.resample_nchw(x4, 0, 0)
        else:
            K.colsum(da, torch.zeros(2 * L, device=DEV))
elif case.startswith("wgg_"):
    # grouped weight gradients of one coupling block (what the training step launches): the four convolutions of a 3x3 or
    # 1x1 GLOW block at level 0 / level 1, bias gradients included
    lvl, taps = {"wgg_l0_3x3": (0, 9), "wgg_l1_3x3": (1, 9), "wgg_l0_1x1": (0, 1), "wgg_l1_1x1": (1, 1)}[case]
    hw, c = (64, 48) if lvl == 0 else (32, 192)
    npix, k = B * hw * hw, 3 if taps == 9 else 1
    jobs = []
    for _ in range(2):
        for cin, cout in ((256, c), (c // 2, 256)):
            x = torch.randn(npix, cin, device=DEV).to(torch.bfloat16)
            dy = torch.randn(npix, cout, device=DEV).to(torch.bfloat16)
            jobs.append((x, dy, (B, hw, hw), taps, torch.zeros(cout, cin, k, k, device=DEV), True, torch.zeros(cout, device=DEV), True))
    for _ in range(3):
        K.wgrad_group(jobs)
elif case.startswith("wg_"):
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
