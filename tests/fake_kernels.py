"""TEST-ONLY stand-in for sin_inn_b200.kernels: the same call signatures implemented with plain
torch ops on CPU tensors.  It lets the `-m "not gpu"` suite exercise the host-side logic (plan
compilation, in-place half-steps, permutation folding, backward-by-inverse bookkeeping, gradient
plumbing) where no GPU exists.  It is never imported by the package; GPU tests use the real library.
"""
import torch
import torch.nn.functional as F

GLOW, IRN = 0, 1


def _haar_fwd(a, b, c, d, scale):
    return ((a + b + c + d) * scale, (a - b + c - d) * scale, (a + b - c - d) * scale, (a - b - c + d) * scale)


def resample_nchw(x, mode, rev, scale=1.0):
    B, c, h, w = x.shape
    if not rev:
        a, b = x[:, :, 0::2, 0::2], x[:, :, 0::2, 1::2]
        cc, d = x[:, :, 1::2, 0::2], x[:, :, 1::2, 1::2]
        o = _haar_fwd(a, b, cc, d, scale) if mode == 1 else (a, b, cc, d)
        return torch.cat(o, 1).contiguous()
    C = c // 4
    o = [x[:, k * C:(k + 1) * C] for k in range(4)]
    if mode == 1:
        a, b, cc, d = _haar_fwd(o[0], o[1], o[2], o[3], scale)
    else:
        a, b, cc, d = o
    out = x.new_empty(B, C, 2 * h, 2 * w)
    out[:, :, 0::2, 0::2], out[:, :, 0::2, 1::2] = a, b
    out[:, :, 1::2, 0::2], out[:, :, 1::2, 1::2] = cc, d
    return out


def resample_nhwc(x, mode, rev, scale=1.0):
    return resample_nchw(x.permute(0, 3, 1, 2), mode, rev, scale).permute(0, 2, 3, 1).contiguous()


def _bf(t2d, rng):
    if rng is None:
        return None
    return t2d[:, rng[0]:rng[1]].to(torch.bfloat16).contiguous()


def nchw_to_nhwc(x, chan_map=None, bf16_range=None):
    if chan_map is not None:
        x = x[:, chan_map.long()]
    out = x.permute(0, 2, 3, 1).contiguous()
    return out, _bf(out.view(-1, out.shape[3]), bf16_range)


def nhwc_to_nchw(x, chan_map=None):
    if chan_map is not None:
        x = x[..., chan_map.long()]
    return x.permute(0, 3, 1, 2).contiguous()


def squeeze2_to_nhwc(x, bf16_range=None):
    return nchw_to_nhwc(resample_nchw(resample_nchw(x, 0, 0), 0, 0), None, bf16_range)


def nhwc_to_unsqueeze2(x):
    return resample_nchw(resample_nchw(nhwc_to_nchw(x, None), 0, 1), 0, 1)


def permute_nhwc(x, chan_map, bf16_range=None):
    out = x[..., chan_map.long()].contiguous()
    return out, _bf(out.view(-1, out.shape[-1]), bf16_range)


def gather_windows_u8(video, centers, win, crop=None):
    T, H, W, Cc = video.shape
    y0, x0, ph, pw = crop if crop is not None else (0, 0, H, W)
    outs = []
    for c in centers.tolist():
        fr = [video[min(max(t, 0), T - 1), y0:y0 + ph, x0:x0 + pw] for t in range(c - win, c + win + 1)]
        outs.append(torch.cat(fr, dim=-1).permute(2, 0, 1).float() / 255.)
    return torch.stack(outs)


def quantize_u8_hwc(x):
    return (x.clamp(0, 1) * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def permute_nhwc_pair(xa, xb, chan_map, bf16_range=None):
    oa, bf = permute_nhwc(xa, chan_map, bf16_range)
    return oa, permute_nhwc(xb, chan_map)[0], bf


def _log_scale(kind, clamp, raw):
    if kind == GLOW:
        r = raw / clamp
        return clamp * 0.636 * torch.atan(r), 0.636 / (1 + r * r)
    sg = torch.sigmoid(raw)
    return clamp * (2 * sg - 1), 2 * clamp * sg * (1 - sg)


def coupling_apply(u, s, t, kind, clamp, inverse, want_bf16=False, fast=False):
    g, _ = _log_scale(kind, clamp, s)
    e = torch.exp(g)
    u.copy_((u - t) / e if inverse else e * u + t)
    return u.to(torch.bfloat16).contiguous() if want_bf16 else None


def coupling_bwd(u, du, s, t, kind, clamp, inverse, ds, dt, want_bf16=False, fast=False):
    g, dg = _log_scale(kind, clamp, s)
    e = torch.exp(g)
    y, dy = u.clone(), du.clone()
    if not inverse:
        x = (y - t) / e
        du.copy_(dy * e)
        ds.copy_((dy * x * e * dg).to(ds.dtype))
        dt.copy_(dy.to(dt.dtype))
    else:
        x = y * e + t
        du.copy_(dy / e)
        ds.copy_((-dy * y * dg).to(ds.dtype))
        dt.copy_((-dy / e).to(dt.dtype))
    u.copy_(x)
    return x.to(torch.bfloat16).contiguous() if want_bf16 else None


def coupling_apply_permute(U, chan_map, rng, s, t, kind, clamp, inverse, bf16_range=None, fast=False):
    C_ = U.shape[-1]
    W = U.reshape(-1, C_).clone()
    coupling_apply(W[:, rng[0]:rng[1]], s, t, kind, clamp, inverse)
    return permute_nhwc(W.reshape(U.shape), chan_map, bf16_range)


def coupling_bwd_unpermute(Y, dY, chan_map, rng, s, t, kind, clamp, inverse, ds, dt, fast=False):
    X, dX, _ = permute_nhwc_pair(Y, dY, chan_map, None)
    C_ = Y.shape[-1]
    x2, dx2 = X.reshape(-1, C_), dX.reshape(-1, C_)
    coupling_bwd(x2[:, rng[0]:rng[1]], dx2[:, rng[0]:rng[1]], s, t, kind, clamp, inverse, ds, dt)
    return X, dX


def cast_slice(src, out, scale=1.0):
    out.copy_((src * scale).to(out.dtype))
    return out


def act_bwd(d, y, out, act, slope=0.0):
    yf = y.float()
    g = torch.where(yf > 0, torch.ones_like(yf), torch.full_like(yf, slope if act == 2 else 0.0)) if act else 1.0
    out.copy_((d.float() * g).to(out.dtype))
    return out


def colsum(src, out, accumulate=False):
    s = src.float().sum(0)
    out.copy_(out + s if accumulate else s)
    return out


def axpy_slice(out, a, alpha):
    out.add_(a.float() * alpha)


SPLIT_BLOCKS = 6


def pack_weight(w, mode, dtype, rows_pad, k_pad):
    co, ci, kh, kw = w.shape
    taps = kh * kw
    wt = w.detach().reshape(co, ci, taps)
    if mode in (2, 3):                   # split packs: K = six blocks [wh | wh | wm | wh | wm | wl]
        kp = k_pad // SPLIT_BLOCKS
        base = pack_weight(w, mode - 2, torch.float32, rows_pad, kp)
        wh = base.to(torch.bfloat16).float()
        wm = (base - wh).to(torch.bfloat16).float()
        return torch.cat((wh, wh, wm, wh, wm, base - wh - wm), dim=2).to(torch.bfloat16)
    out = torch.zeros(taps, rows_pad, k_pad, dtype=torch.float32, device=w.device)
    if mode == 4:                        # fprop, rows interleaved (s_0, t_0, s_1, t_1, ...)
        base = wt.permute(2, 0, 1)
        out[:, 0:co:2, :ci] = base[:, :co // 2]
        out[:, 1:co:2, :ci] = base[:, co // 2:]
    elif mode == 0:
        out[:, :co, :ci] = wt.permute(2, 0, 1)
    else:
        out[:, :ci, :co] = wt.flip(2).permute(2, 1, 0)
    return out.to(dtype)


def split_bf16(src, scale=1.0, blocks=None):
    npix, L = src.shape
    Lp = (L + 7) // 8 * 8
    x = torch.zeros(npix, Lp, dtype=torch.float32, device=src.device)
    x[:, :L] = src.float() * scale
    h = x.to(torch.bfloat16).float()
    m = (x - h).to(torch.bfloat16).float()
    l = x - h - m
    return torch.cat((h, m, h, l, m, h) if blocks in (None, 6) else (h, m, h, l), dim=1).to(torch.bfloat16)


def pack_weights_batched(jobs, njobs, dtype):
    raise NotImplementedError("the CPU stand-in packs weights one by one")


def _shift(x4, dy, dx):
    """x4: [B,H,W,C]; result[b,h,w] = x4[b,h+dy,w+dx] with zero fill."""
    B, H, W, C = x4.shape
    xp = F.pad(x4, (0, 0, 1, 1, 1, 1))
    return xp[:, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]


# ReLU sign-bit masks: element e of a 32-element word sits at bit (e >> 1) + 16 * (e & 1) (include/sininn.h: the two halves of
# packed bf16x2 register j map to bits j and 16 + j)
_BIT_POS = torch.tensor([(e >> 1) + 16 * (e & 1) for e in range(32)])


def _unpack_bits(bits, cout):
    w = bits.to(torch.int64) & 0xffffffff
    return ((w.unsqueeze(-1) >> _BIT_POS) & 1).reshape(bits.shape[0], -1)[:, :cout].float()


def _pack_bits(vals):
    npix, c = vals.shape
    words = (c + 31) // 32
    b = torch.zeros(npix, words * 32, dtype=torch.int64)
    b[:, :c] = (vals > 0).to(torch.int64)
    w = (b.reshape(npix, words, 32) << _BIT_POS).sum(-1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32)


def conv(x, wpack, geom, cout, out, bias=None, act=0, slope=0.0, mask=None, mask_act=0, accumulate=False, alpha=1.0,
         tensor_core=False, mask_bits=None, bits_out=None, coupling=None):
    if coupling is not None:            # interleaved pack (mode 4) + GLOW half-step in the "epilogue"
        L = cout // 2
        a = torch.empty(x.shape[0], cout, dtype=torch.float32, device=x.device)
        il_bias = None
        if bias is not None:
            il_bias = torch.empty(cout, dtype=torch.float32, device=x.device)
            il_bias[0::2], il_bias[1::2] = bias.detach()[:L], bias.detach()[L:]
        conv(x, wpack, geom, cout, a, bias=il_bias, tensor_core=tensor_core)
        s_, t_ = a[:, 0::2].contiguous(), a[:, 1::2].contiguous()
        u = coupling["u"]
        if coupling["mode"] == 1:
            bf = coupling_apply(u, s_, t_, 0, coupling["clamp"], coupling["inverse"], coupling.get("bf16") is not None)
            if coupling.get("a") is not None:
                coupling["a"][:, :L], coupling["a"][:, L:] = s_, t_
        else:
            da = coupling["da"]
            bf = coupling_bwd(u, coupling["du"], s_, t_, 0, coupling["clamp"], coupling["inverse"], da[:, :L], da[:, L:],
                              coupling.get("bf16") is not None)
        if coupling.get("bf16") is not None:
            coupling["bf16"].copy_(bf)
        return
    B, H, W = geom
    cin = x.shape[1]
    taps = wpack.shape[0]
    x4 = x.float().reshape(B, H, W, cin)
    acc = torch.zeros(B * H * W, cout, dtype=torch.float32, device=x.device)
    for tap in range(taps):
        dy, dx = (tap // 3 - 1, tap % 3 - 1) if taps == 9 else (0, 0)
        xs = _shift(x4, dy, dx).reshape(B * H * W, cin)
        acc += xs @ wpack[tap, :cout, :cin].float().t()
    if bias is not None:
        acc = acc + bias.detach().float()
    if act == 1:
        acc = torch.relu(acc)
    elif act == 2:
        acc = F.leaky_relu(acc, slope)
    if mask is not None:
        m = mask.float()
        acc = acc * torch.where(m > 0, torch.ones_like(m), torch.full_like(m, slope if mask_act == 2 else 0.0))
    if mask_bits is not None:
        acc = acc * _unpack_bits(mask_bits, cout)
    acc = acc * alpha
    if bits_out is not None:
        bits_out.copy_(_pack_bits(acc))
    if accumulate:
        acc = acc + out.float()
    out.copy_(acc.to(out.dtype))
    return out


def subnet1x1_supported(cin, hidden, cout):
    return hidden % 64 == 0 and 64 <= hidden <= 256 and cout <= 256 and cout % 4 == 0 and cin % 8 == 0


def subnet1x1_fwd(x, w1pack, b1, w2pack, b2, out, h_out=None, bits_out=None, mask_bits=None, accumulate=False, coupling=None):
    if coupling is not None:
        geom = (1, 1, x.shape[0])
        hidden = w1pack.shape[1]
        h = torch.empty(x.shape[0], hidden, dtype=torch.bfloat16, device=x.device)
        bits = torch.empty(x.shape[0], hidden // 32, dtype=torch.int32, device=x.device)
        conv(x, w1pack, geom, hidden, h, bias=b1, act=1, tensor_core=True, bits_out=bits)
        if h_out is not None:
            h_out.copy_(h)
        if bits_out is not None:
            bits_out.copy_(bits)
        conv(h, w2pack, geom, 2 * coupling["u"].shape[1], None, bias=b2, tensor_core=True, coupling=coupling)
        return None
    npix, cin = x.shape
    hidden, cout = w1pack.shape[1], out.shape[1]
    h = x.float() @ w1pack[0, :hidden, :cin].float().t()
    if mask_bits is not None:                 # gradient mode: masked by the forward's ReLU sign bits
        h = h * _unpack_bits(mask_bits, hidden)
    else:
        h = torch.relu(h + (0 if b1 is None else b1.detach().float()))
    hb = h.to(torch.bfloat16)                 # the kernel feeds the second GEMM with the bf16-rounded hidden tile
    if bits_out is not None:
        bits_out.copy_(_pack_bits(hb.float()))
    if h_out is not None:
        h_out.copy_(hb)
    o = hb.float() @ w2pack[0, :cout, :hidden].float().t() + (0 if b2 is None else b2.detach().float())
    out.copy_(out + o if accumulate else o)
    return out


def subnet1x1_bwd_supported(cin, hidden, cout):
    return hidden == 256 and cin % 8 == 0 and cin <= 32 and cout % 4 == 0 and cout <= 64


def subnet1x1_bwd(x, da, w1pack, b1, w2dpack, w1dpack, dsrc, grads1, grads2):
    npix, cin = x.shape
    hidden, cout = w1pack.shape[1], da.shape[1]
    hb = torch.relu(x.float() @ w1pack[0, :hidden, :cin].float().t() + (0 if b1 is None else b1.detach().float())).to(torch.bfloat16)
    dh = (da.float() @ w2dpack[0, :hidden, :cout].float().t()) * (hb.float() > 0)
    dhb = dh.to(torch.bfloat16)
    dsrc.add_(dhb.float() @ w1dpack[0, :cin, :hidden].float().t())
    (dw1, acc1, db1, accb1), (dw2, acc2, db2, accb2) = grads1, grads2
    for g, acc, val in ((dw2, acc2, (da.float().t() @ hb.float()).reshape(dw2.shape)), (db2, accb2, da.float().sum(0)),
                        (dw1, acc1, (dhb.float().t() @ x.float()).reshape(dw1.shape)), (db1, accb1, dhb.float().sum(0))):
        g.copy_(g + val if acc else val)


def wgrad_group(jobs):
    for job in jobs:
        x, dy, geom, taps, dw, acc, dbias, dbacc = job[:8]
        if len(job) > 8 and job[8] is not None:       # split operands: sum over the channel-block pairs
            xo, yo, mask = job[8]
            cin, cout = x.shape[1], dy.shape[1]
            xb, yb = x.data_ptr(), dy.data_ptr()
            xfull = torch.as_strided(x, (x.shape[0], x.stride(0)), (x.stride(0), 1))
            yfull = torch.as_strided(dy, (dy.shape[0], dy.stride(0)), (dy.stride(0), 1))
            for t in range(len(xo)):
                use_b = dbias is not None and (mask >> t) & 1
                wgrad(xfull[:, xo[t]:xo[t] + cin], yfull[:, yo[t]:yo[t] + cout], geom, taps, dw, accumulate=acc or t > 0,
                      tensor_core=True, dbias=dbias if use_b else None, dbias_accumulate=dbacc or (use_b and t > 0))
            continue
        wgrad(x, dy, geom, taps, dw, accumulate=acc, tensor_core=True, dbias=dbias, dbias_accumulate=dbacc)


def wgrad_merged_supported(x, dy, geom, taps):
    return max(x.shape[1], dy.shape[1]) > 128


def wgrad_merged(x, dy, geom, taps, segments):
    for row0, rows, cin, dw, acc, dbias, dbacc in segments:
        wgrad(x[:, :cin], dy[:, row0:row0 + rows], geom, taps, dw, accumulate=acc, tensor_core=True, dbias=dbias, dbias_accumulate=dbacc)


def wgrad(x, dy, geom, taps, dw, accumulate=False, tensor_core=False, dbias=None, dbias_accumulate=False):
    B, H, W = geom
    if dbias is not None:
        bsum = dy.float().sum(0)
        dbias.copy_(dbias + bsum if dbias_accumulate else bsum)
    cin, cout = x.shape[1], dy.shape[1]
    x4 = x.float().reshape(B, H, W, cin)
    dyf = dy.float()
    res = torch.zeros(cout, cin, taps, device=x.device)
    for tap in range(taps):
        oy, ox = (tap // 3 - 1, tap % 3 - 1) if taps == 9 else (0, 0)
        xs = _shift(x4, oy, ox).reshape(B * H * W, cin)
        res[:, :, tap] = dyf.t() @ xs
    res = res.reshape(dw.shape)
    dw.copy_(dw + res if accumulate else res)
    return dw


def sqdiff(a, b, scale, want_grad=False):
    d = a - (b if b is not None else 0)
    return (d * d).sum() * scale, (2 * scale * d if want_grad else None)


def inn_fwd_loss(y, lr, w_rec, w_nll):
    y = y.detach().clone().requires_grad_(True)
    L = lr.shape[1]
    loss = w_rec * torch.mean((y[:, :L] - lr) ** 2)
    if y.shape[1] > L:
        loss = loss + w_nll * torch.mean(y[:, L:] ** 2)
    loss.backward()
    return loss.detach(), y.grad


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, step, grad_scale=1.0):
    g = grad * grad_scale + weight_decay * param
    exp_avg.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    exp_avg_sq.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    param.sub_((lr / bc1) * exp_avg / (exp_avg_sq.sqrt() / bc2 ** 0.5 + eps))


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, step_state, grad_scale=1.0, grad_b=None):
    step_state[0] += 1
    g = grad if grad_b is None else grad + grad_b
    adam_step(param, g, exp_avg, exp_avg_sq, lr, betas, eps, weight_decay, int(step_state[0]), grad_scale)


def latent_to_nhwc(lr, z, z_dims, chan_map=None, bf16_range=None, seed=0, offset=0, temp=1.0, z_out=None, step_state=None):
    if z is None:
        g = torch.Generator().manual_seed(int(seed) + int(offset))
        z = temp * torch.randn(lr.shape[0], z_dims, lr.shape[2], lr.shape[3], generator=g)
        if z_out is not None:
            z_out.copy_(z)
    return nchw_to_nhwc(torch.cat((lr, z), 1), chan_map, bf16_range)


def channel_affine(U, log_scale, bias, inverse):
    if inverse:
        U.copy_((U - bias) * torch.exp(-log_scale))
    else:
        U.copy_(U * torch.exp(log_scale) + bias)
    return U


def channel_affine_bwd(U, dU, log_scale, bias, inverse, dls, dbias, accumulate):
    s = torch.exp(log_scale)
    y, dy = U.clone(), dU.clone()
    red = tuple(range(U.dim() - 1))
    if not inverse:
        U.copy_((y - bias) / s)
        dU.copy_(dy * s)
        a, b = (dy * (y - bias)).sum(red), dy.sum(red)
    else:
        U.copy_(y * s + bias)
        dU.copy_(dy / s)
        a, b = -(dy * y).sum(red), -(dy / s).sum(red)
    if accumulate:
        dls += a
        dbias += b
    else:
        dls.copy_(a)
        dbias.copy_(b)


def logscale_sum(s, B, kind, clamp, sign, out, accumulate):
    if kind == 0:
        g = clamp * 0.636 * torch.atan(s / clamp)
    else:
        g = clamp * (2 * torch.sigmoid(s) - 1)
    v = sign * g.reshape(B, -1).sum(1)
    if accumulate:
        out += v
    else:
        out.copy_(v)
    return out
