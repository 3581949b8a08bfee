"""Smish activation of the LocalStage CNN (models/local_stage.py:4-6) as one fused elementwise kernel, forward and backward
(SURVEY.md section 8f #4).  `SmishFused` is a drop-in for the reference's `Smish` module: the eager expression launches five
elementwise kernels and keeps four intermediates for autograd; this one reads x once and writes y once (8 B/element forward,
12 B/element backward: HBM bound).  `patch_reference_smish(models.local_stage)` swaps the class before LocalStage is built."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib

_CTX = {}


def _ctx(device):
    """The elementwise entry points need a library context (device binding, stream plumbing); geometry is irrelevant here."""
    key = torch.device(device).index or 0
    if key not in _CTX:
        _CTX[key] = _lib.Context(_lib.make_config(H=21, W=21, max_batch=1), torch.device('cuda', key))
    return _CTX[key]


class _Smish(Function):
    @staticmethod
    def forward(ctx, x):
        if not x.is_cuda:
            raise _lib.BlurryEdgesError(f'SmishFused needs a CUDA tensor (got {x.device}); there is no CPU fallback')
        xc = x.to(torch.float32).contiguous()
        y = torch.empty_like(xc)
        be = _ctx(xc.device)
        be.call('be_smish', xc, xc.numel(), y)
        ctx.save_for_backward(xc)
        ctx.be, ctx.dtype = be, x.dtype
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        gx = torch.empty_like(xc)
        ctx.be.call('be_smish_bwd', xc, gy.to(torch.float32).contiguous(), xc.numel(), gx)
        return gx.to(ctx.dtype)


def smish(x):
    return _Smish.apply(x)


class SmishFused(nn.Module):
    """Drop-in for `models.local_stage.Smish`."""

    def forward(self, x):
        return _Smish.apply(x)


def patch_reference_smish(local_stage_module):
    """Replace the `Smish` class of the (imported) reference module `models.local_stage`; LocalStage instances built afterwards use
    the fused kernel.  Returns the original class."""
    orig = local_stage_module.Smish
    local_stage_module.Smish = SmishFused
    return orig
