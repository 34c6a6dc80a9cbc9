"""Per-kernel parity: each libsininn entry point (through the ctypes C-ABI wrappers) against the plain-torch
restatement in tests/fake_kernels.py on the same seeded inputs, including ragged / unaligned shapes."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fake_kernels as FK  # noqa: E402

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def K():
    from sin_inn_b200 import kernels
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return kernels


def rnd(*shape, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dtype)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", [(2, 3, 16, 32), (1, 12, 6, 10), (3, 5, 4, 12), (1, 3, 270, 480)])
def test_resample_nchw(K, mode, shape):
    x = rnd(*shape, seed=1)
    scale = 0.25 if mode else 1.0
    y = K.resample_nchw(x.to(DEV), mode, 0, scale)
    ref = FK.resample_nchw(x, mode, 0, scale)
    if mode == 0:
        assert torch.equal(y.cpu(), ref)                      # pure permutation: bit exact
    else:
        assert (y.cpu() - ref).abs().max() <= 1e-6 * ref.abs().max()
    xr = K.resample_nchw(y, mode, 1, 1.0)
    assert (xr.cpu() - x).abs().max() <= (0 if mode == 0 else 1e-6 * x.abs().max())


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shape", [(2, 8, 12, 48), (1, 6, 10, 12), (2, 4, 4, 5)])
def test_resample_nhwc(K, mode, shape):
    x = rnd(*shape, seed=2)
    scale = 0.25 if mode else 1.0
    y = K.resample_nhwc(x.to(DEV), mode, 0, scale)
    ref = FK.resample_nhwc(x, mode, 0, scale)
    assert (y.cpu() - ref).abs().max() <= 1e-6 * ref.abs().max()
    xr = K.resample_nhwc(y, mode, 1, 1.0)
    assert (xr.cpu() - x).abs().max() <= 1e-6 * x.abs().max()


@pytest.mark.parametrize("shape", [(2, 3, 16, 32), (3, 3, 64, 256), (1, 3, 120, 280), (2, 8, 8, 12), (1, 1, 4, 4)])
def test_two_squeezes_and_layout_change_in_one_pass(K, shape):
    """squeeze, squeeze, NCHW -> NHWC (the SRF entry, archs.py:28-38) as one kernel == the three separate kernels, bit for
    bit, with the bf16 operand copy of a channel range; and the inverse map restores the input."""
    x = rnd(*shape, seed=8).to(DEV)
    C = 16 * shape[1]
    hint = (C // 2, C) if (C // 2) % 8 == 0 else None
    ref, ref_bf = K.nchw_to_nhwc(K.resample_nchw(K.resample_nchw(x, 0, 0), 0, 0), None, hint)
    got, got_bf = K.squeeze2_to_nhwc(x, hint)
    assert got.shape == ref.shape and torch.equal(got, ref)
    if hint:
        assert torch.equal(got_bf, ref_bf)
    assert torch.equal(got.cpu(), FK.squeeze2_to_nhwc(x.cpu(), hint)[0])
    back = K.nhwc_to_unsqueeze2(got)
    assert torch.equal(back, x)
    assert torch.equal(back, K.resample_nchw(K.resample_nchw(K.nhwc_to_nchw(ref, None), 0, 1), 0, 1))


@pytest.mark.parametrize("C,hw", [(48, (8, 8)), (192, (5, 9)), (7, (3, 11)), (48, (20, 13)), (64, (16, 16)), (72, (9, 15))])
def test_layout_and_permute(K, C, hw):
    x = rnd(2, C, *hw, seed=3)
    perm = torch.randperm(C, generator=torch.Generator().manual_seed(4)).to(torch.int32)
    rng = (C // 2, C) if C % 8 == 0 else None
    y, bf = K.nchw_to_nhwc(x.to(DEV), perm.to(DEV), rng)
    ry, rbf = FK.nchw_to_nhwc(x, perm, rng)
    assert torch.equal(y.cpu(), ry)
    if rng:
        assert torch.equal(bf.cpu(), rbf)
    back = K.nhwc_to_nchw(y, None)
    assert torch.equal(back.cpu(), x[:, perm.long()])
    inv = torch.empty_like(perm)
    inv[perm.long()] = torch.arange(C, dtype=torch.int32)
    assert torch.equal(K.nhwc_to_nchw(y, inv.to(DEV)).cpu(), x)          # out channel i <- in channel map[i]
    z, zbf = K.permute_nhwc(y, perm.to(DEV), rng)
    rz, rzbf = FK.permute_nhwc(ry, perm, rng)
    assert torch.equal(z.cpu(), rz)
    if rng:
        assert torch.equal(zbf.cpu(), rzbf)
    # two tensors, one map, one launch (falls back to two launches when C % 4 != 0)
    y2 = (y * 2 + 1).contiguous()
    pa, pb, pbf = K.permute_nhwc_pair(y, y2, perm.to(DEV), rng)
    assert torch.equal(pa.cpu(), rz) and torch.equal(pb.cpu(), FK.permute_nhwc(y2.cpu(), perm)[0])
    if rng:
        assert torch.equal(pbf.cpu(), rzbf)


@pytest.mark.parametrize("kind,clamp", [(0, 1.2), (1, 1.0)])
@pytest.mark.parametrize("npix,C,L", [(1000, 48, 24), (333, 192, 108), (77, 10, 3)])
def test_coupling_apply_and_bwd(K, kind, clamp, npix, C, L):
    U = rnd(npix, C, seed=5)
    A = rnd(npix, 2 * L, seed=6) * 2
    for inverse in (0, 1):
        u_ref = U.clone()
        FK.coupling_apply(u_ref[:, :L], A[:, :L], A[:, L:], kind, clamp, inverse)
        u = U.clone().to(DEV)
        a = A.to(DEV)
        bf = K.coupling_apply(u[:, :L], a[:, :L], a[:, L:], kind, clamp, inverse, want_bf16=True)
        # fp64 ground truth: the GPU must be as close to it as fp32 arithmetic allows (the division by a small
        # exp(g) amplifies rounding, so the bound is relative to the fp32 CPU result's own error, with slack)
        u64 = U.clone().double()
        FK.coupling_apply(u64[:, :L], A[:, :L].double(), A[:, L:].double(), kind, clamp, inverse)
        err_gpu = (u.cpu().double() - u64).abs().max().item()
        err_cpu = (u_ref.double() - u64).abs().max().item()
        assert err_gpu <= max(4 * err_cpu, 4e-6 * u64.abs().max().item()), (err_gpu, err_cpu)
        assert torch.equal(u.cpu()[:, L:], U[:, L:])                    # untouched half
        assert (bf.float().cpu() - u_ref[:, :L]).abs().max() <= 8e-3 * u_ref[:, :L].abs().max()
        # backward from the output restores the input and matches autograd
        x0 = U[:, :L].clone().double().requires_grad_(True)
        s0 = A[:, :L].clone().double().requires_grad_(True)
        t0 = A[:, L:].clone().double().requires_grad_(True)
        g, _ = FK._log_scale(kind, clamp, s0)
        y0 = (x0 - t0) / torch.exp(g) if inverse else torch.exp(g) * x0 + t0
        dy = rnd(npix, L, seed=7)
        y0.backward(dy.double())
        du = torch.zeros(npix, C)
        du[:, :L] = dy
        du = du.to(DEV)
        ds = torch.empty(npix, L, device=DEV)
        dt = torch.empty(npix, L, device=DEV)
        K.coupling_bwd(u[:, :L], du[:, :L], a[:, :L], a[:, L:], kind, clamp, inverse, ds, dt)
        tol = lambda r: 1e-5 * max(1.0, r.abs().max().item())
        assert (u.cpu()[:, :L] - U[:, :L]).abs().max() <= tol(U)
        assert (du.cpu()[:, :L] - x0.grad.float()).abs().max() <= tol(x0.grad)
        assert (ds.cpu() - s0.grad.float()).abs().max() <= tol(s0.grad)
        assert (dt.cpu() - t0.grad.float()).abs().max() <= tol(t0.grad)


def test_small_helpers(K):
    src = rnd(500, 40, seed=8)
    out = torch.empty(500, 24, dtype=torch.bfloat16, device=DEV)
    K.cast_slice(src.to(DEV)[:, 8:32], out, -0.5)
    assert torch.equal(out.cpu(), (src[:, 8:32] * -0.5).to(torch.bfloat16))
    d, y = rnd(500, 40, seed=9), rnd(500, 40, seed=10)
    o = torch.empty(500, 16, dtype=torch.bfloat16, device=DEV)
    K.act_bwd(d.to(DEV)[:, 4:20], y.to(torch.bfloat16).to(DEV)[:, 4:20], o, 2, 0.2)
    ref = torch.empty(500, 16, dtype=torch.bfloat16)
    FK.act_bwd(d[:, 4:20], y.to(torch.bfloat16)[:, 4:20], ref, 2, 0.2)
    assert torch.equal(o.cpu(), ref)
    big = rnd(70000, 50, seed=11)
    cs = torch.empty(50, device=DEV)
    K.colsum(big.to(DEV), cs)
    assert (cs.cpu() - big.double().sum(0).float()).abs().max() <= 1e-4 * big.abs().sum(0).max()
    cs2 = cs.clone()
    K.colsum(big.to(torch.bfloat16).to(DEV)[:, 3:20], cs2[3:20], accumulate=True)
    ref2 = cs.cpu()[3:20] + big.to(torch.bfloat16)[:, 3:20].double().sum(0).float()
    assert (cs2.cpu()[3:20] - ref2).abs().max() <= 1e-3 * ref2.abs().max()
    a = rnd(500, 12, seed=12)
    tgt = src.clone().to(DEV)
    K.axpy_slice(tgt[:, 4:16], a.to(DEV), -1.0)
    assert torch.allclose(tgt.cpu()[:, 4:16], src[:, 4:16] - a) and torch.equal(tgt.cpu()[:, 16:], src[:, 16:])
    loss, grad = K.sqdiff(src.to(DEV), d.to(DEV), 0.01, want_grad=True)
    assert abs(loss.item() - 0.01 * ((src - d) ** 2).sum().item()) <= 1e-4 * abs(loss.item())
    assert torch.allclose(grad.cpu(), 0.02 * (src - d), atol=1e-6)


def test_adam_matches_torch(K):
    p0, g = rnd(10000, seed=13), rnd(10000, seed=14)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.99), weight_decay=1e-5)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in (1, 2, 3):
        ref.grad = g * step
        opt.step()
        K.adam_step(p, (g * step).to(DEV), m, v, 1e-3, (0.9, 0.99), 1e-8, 1e-5, step)
    assert (p.cpu() - ref.detach()).abs().max() < 1e-6


def _conv_case(K, taps, cin, cout, geom, in_dtype, out_dtype, tensor_core, flags=None, seed=20, in_pad=0):
    B, H, W = geom
    npix = B * H * W
    k = 3 if taps == 9 else 1
    w = rnd(cout, cin, k, k, seed=seed) * 0.1
    bias = rnd(cout, seed=seed + 1)
    xw = rnd(npix, cin + in_pad, seed=seed + 2)
    if in_dtype == torch.bfloat16:
        w = w.to(torch.bfloat16).float()
        xw = xw.to(torch.bfloat16).float()
    flags = flags or {}
    rp, kp = (cout + 15) // 16 * 16, (cin + 15) // 16 * 16
    wp_ref = FK.pack_weight(w, 0, in_dtype, rp, kp)
    wp = K.pack_weight(w.to(DEV), 0, in_dtype, rp, kp)
    assert torch.equal(wp.cpu(), wp_ref)
    out_w = cout + 8
    base = rnd(npix, out_w, seed=seed + 3).to(out_dtype)
    mask = rnd(npix, cout, seed=seed + 4).to(out_dtype) if flags.get("mask") else None
    ref = base.clone()
    FK.conv(xw.to(in_dtype)[:, :cin], wp_ref, geom, cout, ref[:, :cout], bias=bias, act=flags.get("act", 0), slope=0.2,
            mask=mask, mask_act=flags.get("mask_act", 0), accumulate=flags.get("accumulate", False),
            alpha=flags.get("alpha", 1.0))
    got = base.clone().to(DEV)
    K.conv(xw.to(in_dtype).to(DEV)[:, :cin], wp, geom, cout, got[:, :cout], bias=bias.to(DEV), act=flags.get("act", 0),
           slope=0.2, mask=None if mask is None else mask.to(DEV), mask_act=flags.get("mask_act", 0),
           accumulate=flags.get("accumulate", False), alpha=flags.get("alpha", 1.0), tensor_core=tensor_core)
    got = got.cpu().float()
    ref = ref.float()
    assert torch.equal(got[:, cout:], base.float()[:, cout:]), "conv wrote outside its channel slice"
    tol = (2e-5 if out_dtype == torch.float32 else 1e-2) * max(1.0, ref.abs().max().item())
    err = (got[:, :cout] - ref[:, :cout]).abs().max().item()
    assert err <= tol, f"conv mismatch {err} > {tol}"


CONV_SHAPES = [(9, 24, 256, (2, 16, 16)), (9, 256, 48, (1, 8, 24)), (1, 24, 256, (2, 8, 8)), (1, 256, 192, (1, 5, 9)),
               (9, 96, 256, (1, 5, 9)), (9, 56, 32, (1, 12, 20)), (9, 20, 10, (2, 3, 5))]


@pytest.mark.parametrize("taps,cin,cout,geom", CONV_SHAPES)
@pytest.mark.parametrize("dt", ["fp32", "bf16"])
def test_conv_simt(K, taps, cin, cout, geom, dt):
    idt = torch.float32 if dt == "fp32" else torch.bfloat16
    _conv_case(K, taps, cin, cout, geom, idt, torch.float32, False, in_pad=4)
    _conv_case(K, taps, cin, cout, geom, idt, idt, False, {"act": 1})
    _conv_case(K, taps, cin, cout, geom, idt, idt, False, {"mask": True, "mask_act": 1})
    _conv_case(K, taps, cin, cout, geom, idt, torch.float32, False, {"accumulate": True, "alpha": -1.0, "act": 2})


TC_SHAPES = [(9, 24, 256, (2, 16, 16)), (9, 256, 48, (1, 8, 24)), (1, 24, 256, (2, 8, 8)), (1, 256, 192, (1, 5, 9)),
             (9, 96, 256, (1, 5, 9)), (9, 256, 192, (2, 8, 32)), (9, 48, 256, (1, 17, 33)), (1, 192, 256, (1, 4, 4)),
             (9, 56, 32, (1, 12, 20)), (9, 152, 24, (1, 9, 7)), (9, 32, 400, (1, 8, 16))]


@pytest.mark.parametrize("taps,cin,cout,geom", TC_SHAPES)
def test_conv_tc(K, taps, cin, cout, geom):
    """tcgen05 path against the torch restatement: every epilogue variant the engine uses."""
    bf = torch.bfloat16
    _conv_case(K, taps, cin, cout, geom, bf, torch.float32, True)
    _conv_case(K, taps, cin, cout, geom, bf, bf, True, {"act": 1})
    _conv_case(K, taps, cin, cout, geom, bf, bf, True, {"mask": True, "mask_act": 1})
    _conv_case(K, taps, cin, cout, geom, bf, torch.float32, True, {"accumulate": True, "alpha": -1.0, "act": 2})


@pytest.mark.parametrize("taps,cin,cout,geom", [(9, 24, 256, (2, 16, 16)), (1, 96, 256, (1, 5, 9)), (9, 48, 256, (1, 17, 33)),
                                               (9, 192, 256, (2, 8, 20)), (9, 32, 72, (1, 7, 7))])
def test_conv_tc_sign_bits(K, taps, cin, cout, geom):
    """fprop emits the ReLU sign bits of its output; the masked data-gradient consumes them (1 bit / element)."""
    B, H, W = geom
    npix = B * H * W
    k = 3 if taps == 9 else 1
    bf = torch.bfloat16
    x = rnd(npix, cin, seed=60).to(bf)
    w = (rnd(cout, cin, k, k, seed=61) * 0.1)
    bias = rnd(cout, seed=62)
    rp, kp = (cout + 15) // 16 * 16, (cin + 15) // 16 * 16
    wp = K.pack_weight(w.to(DEV), 0, bf, rp, kp)
    words = (cout + 31) // 32
    h = torch.empty(npix, cout, dtype=bf, device=DEV)
    bits = torch.zeros(npix, words, dtype=torch.int32, device=DEV)
    K.conv(x.to(DEV), wp, geom, cout, h, bias=bias.to(DEV), act=1, tensor_core=True, bits_out=bits)
    href = torch.empty(npix, cout, dtype=bf)
    FK.conv(x, FK.pack_weight(w, 0, bf, rp, kp), geom, cout, href, bias=bias, act=1)
    got_bits = FK._unpack_bits(bits.cpu(), cout)
    # bits must agree with the kernel's own output (ties at exactly 0 cannot disagree: bit = value > 0)
    assert torch.equal(got_bits, (h.cpu().float() > 0).float())
    assert (h.cpu().float() - href.float()).abs().max() <= 1e-2 * max(1.0, href.float().abs().max().item())
    # masked dgrad: dy [npix, cin2] -> [npix, cout] zeroed where h == 0
    cin2 = 48
    dy = rnd(npix, cin2, seed=63).to(bf)
    w2 = rnd(cin2, cout, k, k, seed=64) * 0.1
    wd = K.pack_weight(w2.to(DEV), 1, bf, rp, (cin2 + 15) // 16 * 16)
    dh = torch.empty(npix, cout, dtype=bf, device=DEV)
    K.conv(dy.to(DEV), wd, geom, cout, dh, mask_bits=bits, tensor_core=True)
    ref = torch.empty(npix, cout, dtype=bf)
    FK.conv(dy, FK.pack_weight(w2, 1, bf, rp, (cin2 + 15) // 16 * 16), geom, cout, ref, mask=h.cpu(), mask_act=1)
    assert (dh.cpu().float() - ref.float()).abs().max() <= 1e-2 * max(1.0, ref.float().abs().max().item())
    assert torch.equal(dh.cpu().float() == 0, (ref.float() == 0) | (dh.cpu().float() == 0))
    assert bool(((h.cpu().float() == 0) <= (dh.cpu().float() == 0)).all())


def test_conv_tc_matches_simt_bitwise_inputs(K):
    """Same bf16 operands through the CUDA-core and the tensor-core kernels agree to fp32 accumulation noise."""
    B, H, W, cin, cout = 4, 64, 64, 256, 48
    x = rnd(B * H * W, cin, seed=40).to(torch.bfloat16).to(DEV)
    w = (rnd(cout, cin, 3, 3, seed=41) * 0.05).to(DEV)
    wp = K.pack_weight(w, 0, torch.bfloat16, 48, 256)
    o1 = torch.empty(B * H * W, cout, device=DEV)
    o2 = torch.empty_like(o1)
    K.conv(x, wp, (B, H, W), cout, o1, tensor_core=False)
    K.conv(x, wp, (B, H, W), cout, o2, tensor_core=True)
    assert (o1 - o2).abs().max().item() <= 1e-4 * o1.abs().max().item()


@pytest.mark.parametrize("taps,cin,cout,geom", [(9, 256, 48, (2, 8, 8)), (9, 24, 256, (1, 5, 9)), (1, 96, 256, (2, 6, 6)),
                                               (9, 152, 32, (1, 7, 5)), (1, 10, 6, (1, 4, 4))])
@pytest.mark.parametrize("dt", ["fp32", "bf16"])
def test_wgrad_and_dgrad_pack_simt(K, taps, cin, cout, geom, dt):
    B, H, W = geom
    npix = B * H * W
    idt = torch.float32 if dt == "fp32" else torch.bfloat16
    k = 3 if taps == 9 else 1
    x = rnd(npix, cin + 8, seed=30).to(idt)
    dy = rnd(npix, cout, seed=31).to(idt)
    dw0 = rnd(cout, cin, k, k, seed=32)
    ref = dw0.clone()
    FK.wgrad(x[:, :cin], dy, geom, taps, ref, accumulate=True)
    got = dw0.clone().to(DEV)
    K.wgrad(x.to(DEV)[:, :cin], dy.to(DEV), geom, taps, got, accumulate=True)
    assert (got.cpu() - ref).abs().max() <= 2e-5 * max(1.0, ref.abs().max().item())
    # wgrad is the adjoint of conv: <dy, conv(x, w)> == <w, wgrad(x, dy)>
    w = rnd(cout, cin, k, k, seed=33)
    rp, kp = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
    wd = K.pack_weight(w.to(DEV), 1, idt, rp, kp)
    assert torch.equal(wd.cpu(), FK.pack_weight(w, 1, idt, rp, kp))
    dx = torch.empty(npix, cin, device=DEV)
    K.conv(dy.to(DEV), wd, geom, cin, dx)
    xr = x[:, :cin].float().reshape(B, H, W, cin).permute(0, 3, 1, 2).clone().requires_grad_(True)
    wq = w.to(idt).float()
    yr = torch.nn.functional.conv2d(xr, wq, padding=k // 2)
    yr.backward(dy.float().reshape(B, H, W, cout).permute(0, 3, 1, 2))
    ref_dx = xr.grad.permute(0, 2, 3, 1).reshape(npix, cin)
    assert (dx.cpu() - ref_dx).abs().max() <= 2e-5 * max(1.0, ref_dx.abs().max().item())


def test_bad_arguments_are_reported(K):
    from sin_inn_b200._lib import SininnError
    with pytest.raises(SininnError):
        K.resample_nchw(torch.zeros(1, 3, 5, 8, device=DEV), 0, 0)          # odd height
    with pytest.raises(SininnError):
        K.resample_nchw(torch.zeros(1, 3, 4, 8), 0, 0)                      # CPU tensor
    with pytest.raises(SininnError):
        K.coupling_apply(torch.zeros(4, 4, device=DEV), torch.zeros(4, 4, device=DEV), torch.zeros(4, 4, device=DEV), 7, 1.0, 0)


WG_SHAPES = [(9, 24, 256, (2, 16, 16)), (9, 256, 48, (2, 16, 16)), (1, 24, 256, (2, 8, 8)), (1, 256, 192, (1, 5, 9)),
             (9, 96, 256, (1, 12, 20)), (9, 256, 192, (1, 16, 32)), (9, 152, 32, (1, 9, 7)), (9, 236, 108, (1, 6, 10)),
             (9, 56, 32, (3, 7, 33)), (1, 96, 256, (2, 6, 6)), (9, 512, 192, (1, 8, 8)), (9, 96, 512, (1, 8, 8))]


@pytest.mark.parametrize("taps,cin,cout,geom", WG_SHAPES)
def test_wgrad_tc(K, taps, cin, cout, geom):
    B, H, W = geom
    npix = B * H * W
    k = 3 if taps == 9 else 1
    bf = torch.bfloat16
    x = rnd(npix, (cin + 15) // 8 * 8, seed=50).to(bf)        # TMA: pixel stride must be a multiple of 8 channels
    dy = rnd(npix, (cout + 15) // 8 * 8, seed=51).to(bf)
    dw0 = rnd(cout, cin, k, k, seed=52)
    ref = dw0.clone()
    FK.wgrad(x[:, :cin], dy[:, :cout], geom, taps, ref, accumulate=True)
    got = dw0.clone().to(DEV)
    K.wgrad(x.to(DEV)[:, :cin], dy.to(DEV)[:, :cout], geom, taps, got, accumulate=True, tensor_core=True)
    err = (got.cpu() - ref).abs().max().item()
    assert err <= 2e-5 * max(1.0, ref.abs().max().item()) * max(1.0, (npix / 256) ** 0.5), err
    got2 = torch.empty_like(got)
    K.wgrad(x.to(DEV)[:, :cin], dy.to(DEV)[:, :cout], geom, taps, got2, accumulate=False, tensor_core=True)
    got3 = torch.empty_like(got)
    K.wgrad(x.to(DEV)[:, :cin], dy.to(DEV)[:, :cout], geom, taps, got3, accumulate=False, tensor_core=True)
    assert torch.equal(got2, got3)                       # deterministic
    # fused bias gradient: column sums of dy from the same two launches, with and without accumulation
    bsum = dy[:, :cout].float().sum(0)
    db0 = rnd(cout, seed=53)
    db = db0.clone().to(DEV)
    got4 = torch.empty_like(got)
    K.wgrad(x.to(DEV)[:, :cin], dy.to(DEV)[:, :cout], geom, taps, got4, tensor_core=True, dbias=db, dbias_accumulate=True)
    assert torch.equal(got4, got2)
    tolb = 1e-5 * max(1.0, bsum.abs().max().item()) * max(1.0, (npix / 256) ** 0.5)
    assert (db.cpu() - (db0 + bsum)).abs().max().item() <= tolb
    db2 = torch.full((cout,), 7.0, device=DEV)
    K.wgrad(x.to(DEV)[:, :cin], dy.to(DEV)[:, :cout], geom, taps, got4, tensor_core=True, dbias=db2)
    assert (db2.cpu() - bsum).abs().max().item() <= tolb


@pytest.mark.parametrize("taps,cin,cout,geom", [(9, 24, 256, (2, 16, 16)), (9, 256, 192, (1, 16, 32)), (1, 96, 256, (2, 6, 6)), (1, 256, 48, (1, 5, 9)),
                                                (9, 48, 256, (1, 17, 33))])
def test_split_operand_conv_is_fp32_accurate(K, taps, cin, cout, geom):
    """fp32-accurate tensor-core convolution: bf16 hi/mid/lo split operands (split_bf16) x split weight packs (modes
    2 / 3) through the ordinary tcgen05 kernels, against an fp64 evaluation: the error must be at fp32 level (the plain
    bf16 path sits at ~3e-3), fprop and dgrad layouts, plus the split weight gradient (two-term blocks + combine)."""
    B, H, W = geom
    npix = B * H * W
    k = 3 if taps == 9 else 1
    x = rnd(npix, cin, seed=21)
    w = rnd(cout, cin, k, k, seed=22) * 0.05
    bias = rnd(cout, seed=23)
    SB = K.SPLIT_BLOCKS
    xs = K.split_bf16(x.to(DEV))
    assert xs.shape == (npix, SB * ((cin + 7) // 8 * 8))
    wp = K.pack_weight(w.to(DEV), 2, torch.bfloat16, (cout + 15) // 16 * 16, SB * ((cin + 7) // 8 * 8))
    out = torch.empty(npix, cout, device=DEV)
    K.conv(xs, wp, geom, cout, out, bias=bias.to(DEV), tensor_core=True)
    import torch.nn.functional as F

    def nchw(t, c):
        return t.double().view(B, H, W, c).permute(0, 3, 1, 2)

    def flat(t, c):
        return t.permute(0, 2, 3, 1).reshape(npix, c)

    ref = flat(F.conv2d(nchw(x, cin), w.double(), bias.double(), padding=k // 2), cout)
    err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    # the TMEM accumulators do not round to nearest: the error grows by ~6e-8 per accumulated MMA step (measured
    # 2.5e-5 over the 864 steps of a 256-channel 3x3 conv), so the bound scales with the step count
    steps = taps * ((SB * ((cin + 7) // 8 * 8) + 63) // 64) * 4
    assert err <= 2e-6 + 6e-8 * steps, err
    # four-block form (two-term weights): the first four channel blocks against the same pack, 2^-17 accurate
    xs4 = K.split_bf16(x.to(DEV), blocks=4)
    assert torch.equal(xs4, xs[:, :xs4.shape[1]])
    out4 = torch.empty(npix, cout, device=DEV)
    K.conv(xs4, wp, geom, cout, out4, bias=bias.to(DEV), tensor_core=True)
    assert (out4.cpu().double() - ref).abs().max().item() / ref.abs().max().item() <= 2e-5 + 6e-8 * steps
    # data gradient: dx = conv(dy, dgrad pack)
    dy = rnd(npix, cout, seed=24) * 1e-3
    dys = K.split_bf16(dy.to(DEV))
    wpd = K.pack_weight(w.to(DEV), 3, torch.bfloat16, (cin + 15) // 16 * 16, SB * ((cout + 7) // 8 * 8))
    dx = torch.zeros(npix, cin, device=DEV)
    K.conv(dys, wpd, geom, cin, dx, tensor_core=True)
    refd = flat(F.conv_transpose2d(nchw(dy, cout), w.double(), padding=k // 2), cin)
    errd = (dx.cpu().double() - refd).abs().max().item() / refd.abs().max().item()
    assert errd <= 2e-6 + 6e-8 * taps * ((SB * ((cout + 7) // 8 * 8) + 63) // 64) * 4, errd
    # weight + bias gradient: the pixels are walked once per product term of the three-term expansions
    cinp, coutp = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
    xo = [0, 0, cinp, cinp, 0, 3 * cinp]
    yo = [0, coutp, 0, coutp, 3 * coutp, 0]
    dw = torch.full((cout, cin, k, k), 0.5, device=DEV)
    db = torch.full((cout,), 0.25, device=DEV)
    if max(cin, cout) > 128:              # (term lists are a CTA-pair kernel feature)
        K.wgrad_group([(xs[:, :cin], dys[:, :cout], geom, taps, dw, True, db, True, (xo, yo, 0b010011))])
        refw = torch.nn.grad.conv2d_weight(nchw(x, cin), (cout, cin, k, k), nchw(dy, cout), padding=k // 2)
        errw = (dw.cpu().double() - 0.5 - refw).abs().max().item() / refw.abs().max().item()
        errb = (db.cpu().double() - 0.25 - dy.double().sum(0)).abs().max().item() / dy.double().sum(0).abs().max().item()
        assert errw <= 1e-4 and errb <= 1e-4, (errw, errb)          # (0.5 / 0.25 offsets cost a few fp32 ulps of the sum)
        dw2, db2 = torch.empty_like(dw), torch.empty_like(db)
        K.wgrad_group([(xs[:, :cin], dys[:, :cout], geom, taps, dw2, False, db2, False, (xo, yo, 0b010011))])
        errw = (dw2.cpu().double() - refw).abs().max().item() / refw.abs().max().item()
        errb = (db2.cpu().double() - dy.double().sum(0)).abs().max().item() / dy.double().sum(0).abs().max().item()
        assert errw <= 6e-6 and errb <= 6e-6, (errw, errb)              # (6x the accumulation steps of one walk)


@pytest.mark.parametrize("group", [[(9, 24, 256), (9, 256, 48)], [(1, 96, 256), (1, 256, 192), (9, 96, 256), (9, 256, 192)],
                                   [(9, 56, 32), (9, 256, 48)], [(9, 512, 192), (1, 24, 256), (9, 236, 108)]])
def test_wgrad_tc_group(K, group):
    """Several weight + bias gradients in one pair of launches: same results as the torch restatement, deterministic,
    including a group that holds a problem the pair kernel does not take (falls back to one-by-one launches)."""
    geom = (2, 16, 24)
    npix = geom[0] * geom[1] * geom[2]
    bf = torch.bfloat16
    jobs, refs = [], []
    for i, (taps, cin, cout) in enumerate(group):
        k = 3 if taps == 9 else 1
        x = rnd(npix, (cin + 15) // 8 * 8, seed=60 + i).to(bf)
        dy = rnd(npix, (cout + 15) // 8 * 8, seed=70 + i).to(bf)
        dw0 = rnd(cout, cin, k, k, seed=80 + i)
        db0 = rnd(cout, seed=90 + i)
        ref = dw0.clone()
        FK.wgrad(x[:, :cin], dy[:, :cout], geom, taps, ref, accumulate=True)
        refs.append((ref, db0 + dy[:, :cout].float().sum(0)))
        jobs.append((x.to(DEV)[:, :cin], dy.to(DEV)[:, :cout], geom, taps, dw0.clone().to(DEV), True, db0.clone().to(DEV), True))
    K.wgrad_group(jobs)
    scale = max(1.0, (npix / 256) ** 0.5)
    for (ref, bref), j in zip(refs, jobs):
        assert (j[4].cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()) * scale
        assert (j[6].cpu() - bref).abs().max().item() <= 1e-5 * max(1.0, bref.abs().max().item()) * scale
    again = [(j[0], j[1], j[2], j[3], torch.empty_like(j[4]), False, torch.empty_like(j[6]), False) for j in jobs]
    K.wgrad_group(again)
    third = [(j[0], j[1], j[2], j[3], torch.empty_like(j[4]), False, torch.empty_like(j[6]), False) for j in jobs]
    K.wgrad_group(third)
    for a, b in zip(again, third):
        assert torch.equal(a[4], b[4]) and torch.equal(a[6], b[6])


@pytest.mark.parametrize("cin,cout,geom,accumulate", [(24, 24, (2, 24, 20), False), (84, 108, (1, 16, 24), True), (108, 84, (2, 9, 13), False)])
def test_wgrad_merged_dense_block(K, cin, cout, geom, accumulate):
    """The five weight (+ bias) gradients of a DenseBlock (archs.py:74-95) as ONE tensor-core problem over the shared
    concatenation == five separate problems (torch restatement), incl. accumulation into existing gradients."""
    B, H, W = geom
    npix, gc, bf = B * H * W, 32, torch.bfloat16
    ctot, call = cin + 4 * gc, 4 * gc + cout
    cat = rnd(npix, (ctot + 7) // 8 * 8, seed=51).to(bf)
    gall = rnd(npix, (call + 7) // 8 * 8, seed=52).to(bf)
    segs_ref, segs_dev = [], []
    for j in range(5):
        rows, ci = (gc if j < 4 else cout), cin + gc * j
        dw0, db0 = rnd(rows, ci, 3, 3, seed=60 + j), rnd(rows, seed=70 + j)
        segs_ref.append((gc * j, rows, ci, dw0.clone(), accumulate, db0.clone(), accumulate))
        segs_dev.append((gc * j, rows, ci, dw0.clone().to(DEV), accumulate, db0.clone().to(DEV), accumulate))
    catd, galld = cat.to(DEV), gall.to(DEV)
    assert K.wgrad_merged_supported(catd[:, :ctot], galld[:, :call], geom, 9)
    FK.wgrad_merged(cat[:, :ctot], gall[:, :call], geom, 9, segs_ref)
    K.wgrad_merged(catd[:, :ctot], galld[:, :call], geom, 9, segs_dev)
    torch.cuda.synchronize()
    for (r0, rows, ci, dwr, _, dbr, _), (_, _, _, dwd, _, dbd, _) in zip(segs_ref, segs_dev):
        assert (dwd.cpu() - dwr).abs().max().item() <= 2e-3 * max(1.0, dwr.abs().max().item()), (r0, rows, ci)
        assert (dbd.cpu() - dbr).abs().max().item() <= 2e-3 * max(1.0, dbr.abs().max().item()), (r0, rows)
    # deterministic
    again = [(a, b, c, (torch.zeros_like(d) if not accumulate else segs_ref[i][3].clone().to(DEV) * 0 + d * 0), False, torch.zeros_like(f), False)
             for i, (a, b, c, d, e, f, g) in enumerate(segs_dev)]
    again2 = [(a, b, c, torch.zeros_like(d), False, torch.zeros_like(f), False) for (a, b, c, d, e, f, g) in segs_dev]
    K.wgrad_merged(catd[:, :ctot], galld[:, :call], geom, 9, again)
    K.wgrad_merged(catd[:, :ctot], galld[:, :call], geom, 9, again2)
    assert all(torch.equal(x[3], y[3]) and torch.equal(x[5], y[5]) for x, y in zip(again, again2))


@pytest.mark.parametrize("cin,hidden,cout,npix", [(24, 256, 48, 128), (24, 256, 48, 20000), (96, 256, 192, 45), (96, 256, 192, 33 * 40),
                                                   (8, 64, 16, 300), (64, 128, 256, 700), (40, 192, 100, 129)])
@pytest.mark.parametrize("keep", [False, True])
def test_subnet1x1_fused_forward(K, cin, hidden, cout, npix, keep):
    """Fused conv1x1 -> ReLU -> conv1x1 (hidden tile in shared memory) against the torch restatement, with and
    without the hidden activation / ReLU sign bits stored for the backward pass; input and output are channel
    slices of wider matrices."""
    bf = torch.bfloat16
    assert K.subnet1x1_supported(cin, hidden, cout)
    w1 = rnd(hidden, cin, 1, 1, seed=70) * 0.2
    w2 = rnd(cout, hidden, 1, 1, seed=71) * 0.1
    b1, b2 = rnd(hidden, seed=72) * 0.3, rnd(cout, seed=73)
    xw = rnd(npix, cin + 8, seed=74).to(bf)
    k1p, n2p, hp = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16, (hidden + 15) // 16 * 16
    w1r, w2r = FK.pack_weight(w1, 0, bf, hp, k1p), FK.pack_weight(w2, 0, bf, n2p, hp)
    w1d, w2d = K.pack_weight(w1.to(DEV), 0, bf, hp, k1p), K.pack_weight(w2.to(DEV), 0, bf, n2p, hp)
    base = rnd(npix, cout + 4, seed=75)
    ref = base.clone()
    href = torch.empty(npix, hidden, dtype=bf)
    bref = torch.zeros(npix, hidden // 32, dtype=torch.int32)
    FK.subnet1x1_fwd(xw[:, :cin], w1r, b1, w2r, b2, ref[:, :cout], h_out=href, bits_out=bref)
    got = base.clone().to(DEV)
    h = torch.zeros(npix, hidden, dtype=bf, device=DEV) if keep else None
    bits = torch.zeros(npix, hidden // 32, dtype=torch.int32, device=DEV) if keep else None
    K.subnet1x1_fwd(xw.to(DEV)[:, :cin], w1d, b1.to(DEV), w2d, b2.to(DEV), got[:, :cout], h_out=h, bits_out=bits)
    got = got.cpu()
    assert torch.equal(got[:, cout:], base[:, cout:]), "fused subnet wrote outside its channel slice"
    tol = 1e-2 * max(1.0, ref[:, :cout].abs().max().item())
    assert (got[:, :cout] - ref[:, :cout]).abs().max().item() <= tol
    if keep:
        hc = h.cpu().float()
        assert (hc - href.float()).abs().max().item() <= 1e-2 * max(1.0, href.float().abs().max().item())
        assert torch.equal(FK._unpack_bits(bits.cpu(), hidden), (hc > 0).float())
    # determinism: the inverse pass re-evaluates the subnet on the same bits and must get the same bits back
    again = base.clone().to(DEV)
    K.subnet1x1_fwd(xw.to(DEV)[:, :cin], w1d, b1.to(DEV), w2d, b2.to(DEV), again[:, :cout])
    assert torch.equal(again.cpu(), got)


def test_subnet1x1_support_query(K):
    assert K.subnet1x1_supported(24, 256, 48) and K.subnet1x1_supported(96, 256, 192) and K.subnet1x1_supported(48, 256, 24)
    assert not K.subnet1x1_supported(192, 256, 96)       # W1 (192 x 256) + hidden tile + rings exceed 227 KB
    assert not K.subnet1x1_supported(24, 512, 48) and not K.subnet1x1_supported(20, 256, 48)


@pytest.mark.parametrize("cin,hidden,cout,npix", [(24, 256, 48, 128), (24, 256, 48, 5000), (24, 128, 48, 777)])
def test_subnet1x1_fused_data_gradient(K, cin, hidden, cout, npix):
    """The fused 1x1 pipeline in gradient mode: dh = relu_mask * (W2^T da) (stored for the weight gradient) and
    dsrc += W1^T dh accumulated into a channel slice of a wider fp32 matrix, against the torch restatement."""
    bf = torch.bfloat16
    assert K.subnet1x1_supported(cout, hidden, cin)
    w1 = rnd(hidden, cin, 1, 1, seed=80) * 0.2           # conv1: cin -> hidden
    w2 = rnd(cout, hidden, 1, 1, seed=81) * 0.1          # conv2: hidden -> cout
    da = rnd(npix, cout + 8, seed=82).to(bf)
    hfwd = rnd(npix, hidden, seed=83)
    bits_ref = FK._pack_bits((hfwd > 0).float())
    hp, cop, cip = (hidden + 15) // 16 * 16, (cout + 15) // 16 * 16, (cin + 15) // 16 * 16
    # dgrad packs: conv2 -> [hidden][cout], conv1 -> [cin][hidden]
    w2r, w1r = FK.pack_weight(w2, 1, bf, hp, cop), FK.pack_weight(w1, 1, bf, cip, hp)
    w2d, w1d = K.pack_weight(w2.to(DEV), 1, bf, hp, cop), K.pack_weight(w1.to(DEV), 1, bf, cip, hp)
    base = rnd(npix, cin + 4, seed=84)
    ref = base.clone()
    dh_ref = torch.empty(npix, hidden, dtype=bf)
    FK.subnet1x1_fwd(da[:, :cout], w2r, None, w1r, None, ref[:, :cin], h_out=dh_ref, mask_bits=bits_ref, accumulate=True)
    got = base.clone().to(DEV)
    dh = torch.zeros(npix, hidden, dtype=bf, device=DEV)
    K.subnet1x1_fwd(da.to(DEV)[:, :cout], w2d, None, w1d, None, got[:, :cin], h_out=dh, mask_bits=bits_ref.to(DEV), accumulate=True)
    got = got.cpu()
    assert torch.equal(got[:, cin:], base[:, cin:]), "fused data gradient wrote outside its channel slice"
    assert (got[:, :cin] - ref[:, :cin]).abs().max().item() <= 1e-2 * max(1.0, ref[:, :cin].abs().max().item())
    dhc = dh.cpu().float()
    assert (dhc - dh_ref.float()).abs().max().item() <= 1e-2 * max(1.0, dh_ref.float().abs().max().item())
    assert bool(((hfwd <= 0) <= (dhc == 0)).all())       # exactly zero wherever the forward ReLU was off


@pytest.mark.parametrize("cin,cout,npix", [(24, 48, 128), (24, 48, 5000), (24, 48, 2 * 148 * 128 + 77), (16, 48, 3001), (32, 32, 1000),
                                           (8, 16, 300)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_subnet1x1_fused_backward(K, cin, cout, npix, accumulate):
    """The whole backward pass of a 1x1 subnet in one kernel (hidden activation re-evaluated on chip, input gradient, both
    weight and both bias gradients) against the torch restatement; operands are channel slices of wider matrices."""
    bf, hidden = torch.bfloat16, 256
    assert K.subnet1x1_bwd_supported(cin, hidden, cout)
    w1 = rnd(hidden, cin, 1, 1, seed=90) * 0.2
    w2 = rnd(cout, hidden, 1, 1, seed=91) * 0.1
    b1 = rnd(hidden, seed=92) * 0.3
    xw = rnd(npix, cin + 8, seed=93).to(bf)
    daw = rnd(npix, cout + 8, seed=94).to(bf)
    cip, cop = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
    packs_ref = (FK.pack_weight(w1, 0, bf, hidden, cip), FK.pack_weight(w2, 1, bf, hidden, cop), FK.pack_weight(w1, 1, bf, cip, hidden))
    packs_dev = (K.pack_weight(w1.to(DEV), 0, bf, hidden, cip), K.pack_weight(w2.to(DEV), 1, bf, hidden, cop),
                 K.pack_weight(w1.to(DEV), 1, bf, cip, hidden))
    base = rnd(npix, cin + 4, seed=95)
    g0 = [rnd(hidden, cin, 1, 1, seed=96), rnd(hidden, seed=97), rnd(cout, hidden, 1, 1, seed=98), rnd(cout, seed=99)]
    ref, gref = base.clone(), [g.clone() for g in g0]
    FK.subnet1x1_bwd(xw[:, :cin], daw[:, :cout], packs_ref[0], b1, packs_ref[1], packs_ref[2], ref[:, :cin],
                     (gref[0], accumulate, gref[1], accumulate), (gref[2], accumulate, gref[3], accumulate))
    got, gg = base.clone().to(DEV), [g.clone().to(DEV) for g in g0]
    xd, dad = xw.to(DEV), daw.to(DEV)

    def run(out, grads):
        K.subnet1x1_bwd(xd[:, :cin], dad[:, :cout], packs_dev[0], b1.to(DEV), packs_dev[1], packs_dev[2], out[:, :cin],
                        (grads[0], accumulate, grads[1], accumulate), (grads[2], accumulate, grads[3], accumulate))
    run(got, gg)
    torch.cuda.synchronize()
    gotc = got.cpu()
    assert torch.equal(gotc[:, cin:], base[:, cin:]), "fused backward wrote outside its channel slice"
    assert (gotc[:, :cin] - ref[:, :cin]).abs().max().item() <= 1e-2 * max(1.0, ref[:, :cin].abs().max().item())
    for name, a, b in zip(("dw1", "db1", "dw2", "db2"), gg, gref):
        err = (a.cpu() - b).abs().max().item()
        assert err <= 2e-3 * max(1.0, b.abs().max().item()), (name, err, b.abs().max().item())
    # bit-deterministic: fixed-order reduction of the per-CTA partials, no float atomics
    got2, gg2 = base.clone().to(DEV), [g.clone().to(DEV) for g in g0]
    run(got2, gg2)
    assert torch.equal(got2.cpu(), gotc) and all(torch.equal(a, b) for a, b in zip(gg, gg2))


def test_subnet1x1_backward_support_query(K):
    assert K.subnet1x1_bwd_supported(24, 256, 48) and K.subnet1x1_bwd_supported(8, 256, 16)
    assert not K.subnet1x1_bwd_supported(96, 256, 192)      # the persistent weight-gradient accumulators exceed 512 TMEM columns
    assert not K.subnet1x1_bwd_supported(24, 128, 48) and not K.subnet1x1_bwd_supported(20, 256, 48)


@pytest.mark.parametrize("B,C,L,hw,w_nll", [(3, 48, 12, (16, 16), 0.0), (2, 192, 84, (5, 9), 0.7), (1, 12, 12, (8, 8), 0.3)])
def test_fused_forward_half_loss(K, B, C, L, hw, w_nll):
    """lit_wrapper.py:45-48 (loss.reconstruction on the LR channels + loss.latent_nll on the z channels) and the gradient
    w.r.t. the network output, one pass."""
    y = rnd(B, C, *hw, seed=90)
    lr = rnd(B, L, *hw, seed=91)
    yr = y.clone().requires_grad_(True)
    ref = 1.3 * torch.mean((yr[:, :L] - lr) ** 2)
    if C > L:
        ref = ref + w_nll * torch.mean(yr[:, L:] ** 2)
    ref.backward()
    loss, grad = K.inn_fwd_loss(y.to(DEV), lr.to(DEV), 1.3, w_nll)
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    assert (grad.cpu() - yr.grad).abs().max().item() <= 1e-6 * max(1e-3, yr.grad.abs().max().item())


def test_quantize_u8_hwc_matches_topilimage_arithmetic(K):
    """lit_wrapper.py:117-121: ToPILImage on a float frame is pic.mul(255).byte(); same bytes for in-range values,
    clamping outside [0, 1]."""
    x = torch.rand(3, 3, 37, 53, generator=torch.Generator().manual_seed(5))
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 1.0 / 255, 254.999 / 255])
    got = K.quantize_u8_hwc(x.to(DEV)).cpu()
    ref = x.mul(255).byte().permute(0, 2, 3, 1)
    assert got.shape == (3, 37, 53, 3) and torch.equal(got, ref)
    y = x * 3 - 1                                         # out of range: clamped to 0 / 255
    g2 = K.quantize_u8_hwc(y.to(DEV)).cpu()
    assert torch.equal(g2, FK.quantize_u8_hwc(y))
    from sin_inn_b200 import train
    host = train.frames_to_uint8(x.to(DEV))
    torch.cuda.synchronize()
    assert host.is_pinned() and torch.equal(host, ref)


def test_gather_windows_u8_matches_data_py_arithmetic(K):
    """data.py:31-45: per sample np.concatenate of the (2*lr_window+1) decoded LR frames along channels,
    transpose(-1, 0, 1), / 255 -- here from a uint8 clip resident on the GPU, with an optional patch crop."""
    g = torch.Generator().manual_seed(9)
    T, h, w, C, win = 40, 18, 26, 4, 3
    video = torch.randint(0, 256, (T, h, w, C), generator=g, dtype=torch.uint8)
    centers = torch.tensor([5, 17, 30, 36], dtype=torch.int32)
    ref = torch.stack([torch.from_numpy(__import__("numpy").concatenate([video[t].numpy() for t in range(c - win, c + win + 1)], axis=-1)
                                        .transpose(2, 0, 1).astype("float32")) / 255. for c in centers.tolist()])
    got = K.gather_windows_u8(video.to(DEV), centers.to(DEV), win)
    assert got.shape == (4, (2 * win + 1) * C, h, w) and torch.equal(got.cpu(), ref)
    crop = (3, 5, 8, 16)
    got = K.gather_windows_u8(video.to(DEV), centers.to(DEV), win, crop)
    assert torch.equal(got.cpu(), ref[:, :, 3:11, 5:21])
    assert torch.equal(FK.gather_windows_u8(video, centers, win, crop), ref[:, :, 3:11, 5:21])
    # the batcher: HR frames (win 0) + LR windows for a set of sample ids
    from sin_inn_b200 import train
    import types
    opt = types.SimpleNamespace(lr_window=win, fps=30)
    hr_frames = torch.randint(0, 256, (3, 2 * h, 2 * w, 3), generator=g, dtype=torch.uint8)
    vb = train.VideoBatcher(video.to(DEV), hr_frames.to(DEV), opt, centers=[5, 17, 30])
    hr, lr = vb.batch(torch.tensor([2, 0]), lr_crop=(2, 4, 8, 8))
    assert torch.equal(lr.cpu(), ref[[2, 0]][:, :, 2:10, 4:12])
    assert torch.equal(hr.cpu(), hr_frames[[2, 0]].permute(0, 3, 1, 2).float()[:, :, 4:20, 8:24] / 255.)


@pytest.mark.parametrize("env", [{"SININN_PAIR": "0"}, {"SININN_PAIR": "0", "SININN_HALO": "0"}, {"SININN_WG_PAIR": "0"},
                                 {"SININN_WGRAD_GROUP": "1"}, {"SININN_WG_HALO": "16"}, {"SININN_PDL": "0"}])
def test_kernel_selection_switches(env):
    """The environment switches select the older kernels (single-CTA halo / per-tap 3x3 convolutions, single-CTA weight
    gradients, ungrouped weight gradients, 16-pixel halo rows, no programmatic dependent launch).  They are read once per
    process, so each setting runs the convolution / weight-gradient / one network parity test in a subprocess."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sel = "test_conv_tc or test_wgrad_tc or (test_bf16_path_matches_oracle and SRF-4-2-10-1-72-104)"
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_kernels.py", "tests/test_gpu_parity.py", "-q", "-x", "-m", "gpu",
                        "-k", sel, "-p", "no:cacheprovider"], cwd=root, env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("npix,C,c0,L", [(1000, 48, 24, 24), (1000, 48, 0, 24), (333, 192, 96, 96), (77, 12, 8, 4)])
def test_coupling_folded_with_permutation(K, npix, C, c0, L, fast):
    """The last half-step of a block + the permutation after it in one kernel (value pass), and the permutation's undo + the
    half-step's backward in one kernel (backward pass) == the separate kernels, bit for bit (same arithmetic per element)."""
    bf = torch.bfloat16
    U = rnd(npix, C, seed=41).to(DEV)
    dU = rnd(npix, C, seed=42).to(DEV)
    a = (rnd(npix, 2 * L, seed=43) * 2).to(DEV)
    perm = torch.randperm(C, generator=torch.Generator().manual_seed(44)).to(torch.int32).to(DEV)
    hint = (C // 2, C) if (C // 2) % 4 == 0 else None
    for inverse in (0, 1):
        # value pass
        ref = U.clone()
        K.coupling_apply(ref[:, c0:c0 + L], a[:, :L], a[:, L:], 0, 1.2, inverse, fast=fast)
        ref_p, ref_bf = K.permute_nhwc(ref.view(1, 1, npix, C), perm, hint)
        got, got_bf = K.coupling_apply_permute(U.view(1, 1, npix, C), perm, (c0, c0 + L), a[:, :L], a[:, L:], 0, 1.2, inverse, hint, fast=fast)
        assert torch.equal(got, ref_p)
        if hint:
            assert torch.equal(got_bf, ref_bf)
        # backward pass: (ref_p, dU) are the permuted trunk / gradient; inv undoes the permutation
        inv = torch.empty_like(perm)
        inv[perm.long()] = torch.arange(C, dtype=torch.int32, device=DEV)
        Yp = ref_p.view(npix, C)
        X1, dX1, _ = K.permute_nhwc_pair(Yp.view(1, 1, npix, C), dU.view(1, 1, npix, C), inv, None)
        X1, dX1 = X1.view(npix, C).clone(), dX1.view(npix, C).clone()
        ds1, dt1 = torch.empty(npix, L, dtype=bf, device=DEV), torch.empty(npix, L, dtype=bf, device=DEV)
        K.coupling_bwd(X1[:, c0:c0 + L], dX1[:, c0:c0 + L], a[:, :L], a[:, L:], 0, 1.2, inverse, ds1, dt1, fast=fast)
        ds2, dt2 = torch.empty(npix, L, dtype=bf, device=DEV), torch.empty(npix, L, dtype=bf, device=DEV)
        X2, dX2 = K.coupling_bwd_unpermute(Yp.view(1, 1, npix, C), dU.view(1, 1, npix, C), inv, (c0, c0 + L), a[:, :L], a[:, L:], 0, 1.2, inverse,
                                           ds2, dt2, fast=fast)
        assert torch.equal(X2.view(npix, C), X1) and torch.equal(dX2.view(npix, C), dX1)
        assert torch.equal(ds2, ds1) and torch.equal(dt2, dt1)
        if not fast:
            assert (X1 - U).abs().max().item() <= 1e-5 * max(1.0, U.abs().max().item())      # the block input is back


@pytest.mark.parametrize("kind,clamp", [(0, 1.2), (1, 1.0)])
@pytest.mark.parametrize("npix,C,L", [(1000, 48, 24), (333, 192, 108), (77, 10, 3)])
def test_coupling_fast_math_variants(K, kind, clamp, npix, C, L):
    """fast_math = 1 (what the bf16 path launches: polynomial atan, ex2.approx, approximate division) against the accurate
    kernels: errors stay at fp32 rounding level, two orders of magnitude inside that path's bf16 operand rounding."""
    U = rnd(npix, C, seed=5).to(DEV)
    a = (rnd(npix, 2 * L, seed=6) * 2).to(DEV)
    dU = rnd(npix, C, seed=7).to(DEV)
    rel = lambda x, y: (x.float() - y.float()).abs().max().item() / max(1.0, y.float().abs().max().item())
    for inverse in (0, 1):
        ua, uf = U.clone(), U.clone()
        K.coupling_apply(ua[:, :L], a[:, :L], a[:, L:], kind, clamp, inverse)
        bf = K.coupling_apply(uf[:, :L], a[:, :L], a[:, L:], kind, clamp, inverse, want_bf16=True, fast=True)
        assert rel(uf, ua) <= 4e-6 and torch.equal(uf[:, L:], U[:, L:])
        assert rel(bf, ua[:, :L]) <= 8e-3
        outs = []
        for fast in (False, True):
            y, dy = ua.clone(), dU.clone()
            ds, dt = torch.empty(npix, L, device=DEV), torch.empty(npix, L, device=DEV)
            K.coupling_bwd(y[:, :L], dy[:, :L], a[:, :L], a[:, L:], kind, clamp, inverse, ds, dt, fast=fast)
            outs.append((y, dy, ds, dt))
        for got, ref in zip(outs[1], outs[0]):
            assert rel(got, ref) <= 1e-5


@pytest.mark.parametrize("L,geom", [(24, (2, 16, 16)), (96, (1, 12, 20)), (24, (1, 17, 33)), (96, (3, 7, 5))])
@pytest.mark.parametrize("inverse", [0, 1])
def test_coupling_fused_into_conv_epilogue(K, L, geom, inverse):
    """GLOW half-step (value pass, cpl_mode 1) and its backward (cpl_mode 2) in the epilogue of the subnet's second 3x3
    convolution (interleaved mode-4 pack) == the unfused convolution followed by coupling_apply / coupling_bwd, bit for bit
    on the subnet output (same accumulation order per column) and to fp32 rounding on the coupled values."""
    B, H, W = geom
    npix, hid, cout, C = B * H * W, 256, 2 * L, 2 * L
    bf = torch.bfloat16
    h = rnd(npix, hid, seed=31).abs().to(bf).to(DEV)
    w = (rnd(cout, hid, 3, 3, seed=32) * 0.03).to(DEV)
    bias = rnd(cout, seed=33).to(DEV)
    U0, dU0 = rnd(npix, C, seed=34).to(DEV), (rnd(npix, C, seed=35) * 1e-2).to(DEV)
    wp = K.pack_weight(w, 0, bf, (cout + 15) // 16 * 16, hid)
    wpi = K.pack_weight(w, 4, bf, (cout + 15) // 16 * 16, hid)
    a = torch.empty(npix, cout, device=DEV)
    K.conv(h, wp, geom, cout, a, bias=bias, tensor_core=True)
    for c0 in (0, L):                                  # the slice may be either half of the trunk
        # value pass
        Ur, Uf = U0.clone(), U0.clone()
        bfr = K.coupling_apply(Ur[:, c0:c0 + L], a[:, :L], a[:, L:], 0, 1.2, inverse, True)
        bff = torch.empty(npix, L, dtype=bf, device=DEV)
        af = torch.full((npix, cout), float("nan"), device=DEV)
        K.conv(h, wpi, geom, cout, None, bias=bias, tensor_core=True,
               coupling=dict(mode=1, u=Uf[:, c0:c0 + L], clamp=1.2, inverse=inverse, bf16=bff, a=af))
        assert torch.equal(af, a)                       # the kept subnet output: same accumulation order per column
        assert (Ur - Uf).abs().max().item() <= 2e-6 * Ur.abs().max().item()
        assert torch.equal(Ur[:, :c0], Uf[:, :c0]) and torch.equal(Ur[:, c0 + L:], Uf[:, c0 + L:])     # the other half is untouched
        assert (bfr.float() - bff.float()).abs().max().item() <= 1e-2 * bfr.float().abs().max().item()
        # backward pass
        Ur, Uf, dUr, dUf = U0.clone(), U0.clone(), dU0.clone(), dU0.clone()
        dar = torch.empty(npix, 2 * L, dtype=bf, device=DEV)
        xr = K.coupling_bwd(Ur[:, c0:c0 + L], dUr[:, c0:c0 + L], a[:, :L], a[:, L:], 0, 1.2, inverse, dar[:, :L], dar[:, L:], True)
        daf = torch.empty(npix, 2 * L, dtype=bf, device=DEV)
        xf = torch.empty(npix, L, dtype=bf, device=DEV)
        K.conv(h, wpi, geom, cout, None, bias=bias, tensor_core=True,
               coupling=dict(mode=2, u=Uf[:, c0:c0 + L], clamp=1.2, inverse=inverse, bf16=xf, du=dUf[:, c0:c0 + L], da=daf))
        assert (Ur - Uf).abs().max().item() <= 2e-6 * Ur.abs().max().item()
        assert (dUr - dUf).abs().max().item() <= 2e-6 * dUr.abs().max().item()
        assert (dar.float() - daf.float()).abs().max().item() <= 1e-2 * dar.float().abs().max().item()
        assert (xr.float() - xf.float()).abs().max().item() <= 1e-2 * xr.float().abs().max().item()


@pytest.mark.parametrize("L,npix", [(24, 128), (24, 5000), (96, 45), (96, 33 * 40)])
@pytest.mark.parametrize("inverse", [0, 1])
def test_coupling_fused_into_1x1_subnet_kernel(K, L, npix, inverse):
    """The fused 1x1 subnet kernel with the GLOW half-step (and its backward) in its second epilogue == the same kernel
    writing the subnet output, followed by coupling_apply / coupling_bwd."""
    bf = torch.bfloat16
    cin, hid, cout, C = L, 256, 2 * L, 2 * L
    x = rnd(npix, cin, seed=41).to(bf).to(DEV)
    w1, w2 = (rnd(hid, cin, 1, 1, seed=42) * 0.2).to(DEV), (rnd(cout, hid, 1, 1, seed=43) * 0.05).to(DEV)
    b1, b2 = rnd(hid, seed=44).to(DEV), rnd(cout, seed=45).to(DEV)
    w1p = K.pack_weight(w1, 0, bf, hid, (cin + 15) // 16 * 16)
    w2p = K.pack_weight(w2, 0, bf, (cout + 15) // 16 * 16, hid)
    w2pi = K.pack_weight(w2, 4, bf, (cout + 15) // 16 * 16, hid)
    U0, dU0 = rnd(npix, C, seed=46).to(DEV), (rnd(npix, C, seed=47) * 1e-2).to(DEV)
    a = torch.empty(npix, cout, device=DEV)
    h_ref = torch.empty(npix, hid, dtype=bf, device=DEV)
    bits_ref = torch.empty(npix, hid // 32, dtype=torch.int32, device=DEV)
    K.subnet1x1_fwd(x, w1p, b1, w2p, b2, a, h_out=h_ref, bits_out=bits_ref)
    c0 = L
    Ur, Uf = U0.clone(), U0.clone()
    bfr = K.coupling_apply(Ur[:, c0:c0 + L], a[:, :L], a[:, L:], 0, 1.2, inverse, True)
    bff = torch.empty(npix, L, dtype=bf, device=DEV)
    af = torch.full((npix, cout), float("nan"), device=DEV)
    K.subnet1x1_fwd(x, w1p, b1, w2pi, b2, None, coupling=dict(mode=1, u=Uf[:, c0:c0 + L], clamp=1.2, inverse=inverse, bf16=bff, a=af))
    assert torch.equal(af, a)
    assert (Ur - Uf).abs().max().item() <= 2e-6 * Ur.abs().max().item()
    assert (bfr.float() - bff.float()).abs().max().item() <= 1e-2 * bfr.float().abs().max().item()
    Ur, Uf, dUr, dUf = U0.clone(), U0.clone(), dU0.clone(), dU0.clone()
    dar = torch.empty(npix, 2 * L, dtype=bf, device=DEV)
    K.coupling_bwd(Ur[:, c0:c0 + L], dUr[:, c0:c0 + L], a[:, :L], a[:, L:], 0, 1.2, inverse, dar[:, :L], dar[:, L:], False)
    daf = torch.empty(npix, 2 * L, dtype=bf, device=DEV)
    h2 = torch.empty_like(h_ref)
    bits2 = torch.empty_like(bits_ref)
    K.subnet1x1_fwd(x, w1p, b1, w2pi, b2, None, h_out=h2, bits_out=bits2,
                    coupling=dict(mode=2, u=Uf[:, c0:c0 + L], clamp=1.2, inverse=inverse, du=dUf[:, c0:c0 + L], da=daf))
    assert torch.equal(h2, h_ref) and torch.equal(bits2, bits_ref)
    assert (Ur - Uf).abs().max().item() <= 2e-6 * Ur.abs().max().item()
    assert (dUr - dUf).abs().max().item() <= 2e-6 * dUr.abs().max().item()
    assert (dar.float() - daf.float()).abs().max().item() <= 1e-2 * dar.float().abs().max().item()
