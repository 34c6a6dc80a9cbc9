"""Option namespace and synthetic inputs of the reference's entry point (main.py:9-83), for callers that do not go
through argparse: benchmarks, tools, tests.  (The oracle has its own copy; the product path never imports oracle/.)"""
import types

import torch


def make_opt(scale=4, num_coupling=4, lr_window=10, **kw):
    """main.py:74-75: lr_dims = (2*lr_window+1)*4, z_dims = scale^2*3*4 - lr_dims; loss weights / optimizer settings are
    main.py's defaults (lambda_fwd_rec 1, lambda_bwd_rec 1, the mmd / nll / tcr lambdas 0; Adam 1e-4, betas (0.9, 0.99),
    weight decay 1e-5; temp 0.8)."""
    lr_dims = (2 * lr_window + 1) * 4
    z_dims = scale * scale * 3 * 4 - lr_dims
    d = dict(scale=scale, num_coupling=num_coupling, lr_window=lr_window, lr_dims=lr_dims, z_dims=z_dims,
             lambda_fwd_rec=1.0, lambda_fwd_mmd=0.0, lambda_latent_nll=0.0, lambda_bwd_rec=1.0, lambda_bwd_mmd=0.0,
             lambda_bwd_tcr=0.0, learning_rate=1e-4, adam_betas=(0.9, 0.99), weight_decay=1e-5, temp=0.8,
             architecture="SRF", fps=30, seed=0)
    d.update(kw)
    return types.SimpleNamespace(**d)


def synthetic_batch(opt, batch, height, width, seed=0, dtype=torch.float32, with_z=True):
    """HR, LR ~ U[0,1) (images are /255, data.py:37-38), z ~ N(0,1) (lit_wrapper.py:41); the LR
    grid is HR / (2*scale).  CPU tensors; with_z=False returns z = None (drawn on the device by the step)."""
    g = torch.Generator().manual_seed(seed)
    f = 2 * opt.scale
    hr = torch.rand(batch, 3, height, width, generator=g, dtype=dtype)
    lr = torch.rand(batch, opt.lr_dims, height // f, width // f, generator=g, dtype=dtype)
    z = torch.randn(batch, opt.z_dims, height // f, width // f, generator=g, dtype=dtype) if with_z else None
    return hr, lr, z
