"""autograd wrappers of the method-granularity kernels (csrc/be_ops.cu): one torch.autograd.Function per method of the
reference's helper classes, forward and backward both on this library's kernels."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _lib


def _f32c(t, device):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t, dtype=torch.float32, device=device)
    if t.device.type != 'cuda':
        raise _lib.BlurryEdgesError(f'blurry_edges_b200 ops need CUDA tensors (got {t.device}); there is no CPU fallback')
    return t.to(torch.float32).contiguous()


class Params2Dists(Function):
    """utils/postprocessing_loss.py:43-86.  params [B,8(+),Hp,Wp] or [B,8(+)] -> [B,2,R,R,Hp,Wp] / [B,2,R,R]"""

    @staticmethod
    def forward(ctx, params, be, R):
        p = _f32c(params[:, :8], params.device)
        B, sp = p.shape[0], tuple(p.shape[2:])
        Lsp = 1
        for d in sp:
            Lsp *= d
        out = torch.empty((B, 2, R, R) + sp, device=p.device, dtype=torch.float32)
        be.call('be_params2dists', p, 8, B, Lsp, out)
        ctx.save_for_backward(p)
        ctx.be, ctx.meta, ctx.nch = be, (B, Lsp), params.shape[1]
        return out

    @staticmethod
    def backward(ctx, gd):
        (p,) = ctx.saved_tensors
        B, Lsp = ctx.meta
        gp8 = torch.empty_like(p)
        ctx.be.call('be_params2dists_bwd', p, 8, _f32c(gd, p.device), B, Lsp, gp8)
        if ctx.nch > 8:
            gp = torch.zeros((B, ctx.nch) + tuple(p.shape[2:]), device=p.device, dtype=torch.float32)
            gp[:, :8] = gp8
            return gp, None, None
        return gp8, None, None


class Dists2Indicators(Function):
    """:91-95.  dists [B,2,R,R,*sp], etas [B,2,*sp] -> wedges [B,3,R,R,*sp]"""

    @staticmethod
    def forward(ctx, dists, etas, be):
        d, e = _f32c(dists, dists.device), _f32c(etas, dists.device)
        B, R = d.shape[0], d.shape[2]
        Lsp = d.numel() // (B * 2 * R * R)
        out = torch.empty((B, 3) + tuple(d.shape[2:]), device=d.device, dtype=torch.float32)
        be.call('be_dists2indicators', d, e, B, Lsp, out)
        ctx.save_for_backward(d, e)
        ctx.be, ctx.meta = be, (B, Lsp)
        return out

    @staticmethod
    def backward(ctx, gw):
        d, e = ctx.saved_tensors
        B, Lsp = ctx.meta
        gd, ge = torch.empty_like(d), torch.empty_like(e)
        ctx.be.call('be_dists2indicators_bwd', d, e, _f32c(gw, d.device), B, Lsp, gd, ge)
        return gd, ge, None


class Elementwise(Function):
    """op 0 params2etas (:88-89), 1 normalized_gaussian(x, delta) (:97-98), 2 depth2sigma(depth, rho_prime)"""

    @staticmethod
    def forward(ctx, x, be, op, p0):
        xc = _f32c(x, x.device)
        y = torch.empty_like(xc)
        be.call('be_elementwise', op, xc, float(p0), xc.numel(), y)
        ctx.save_for_backward(xc)
        ctx.be, ctx.op, ctx.p0 = be, op, float(p0)
        return y

    @staticmethod
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        gx = torch.empty_like(xc)
        ctx.be.call('be_elementwise_bwd', ctx.op, xc, _f32c(gy, xc.device), ctx.p0, xc.numel(), gx)
        return gx, None, None, None


class Etas2Depth(Function):
    """utils/depth_etas.py:23-34"""

    @staticmethod
    def forward(ctx, e1, e2, be):
        a, b = torch.broadcast_tensors(e1, e2)
        a, b = _f32c(a, e1.device), _f32c(b, e1.device)
        z = torch.empty_like(a)
        be.call('be_etas2depth', a, b, a.numel(), z)
        ctx.save_for_backward(a, b)
        ctx.be = be
        return z

    @staticmethod
    def backward(ctx, gz):
        a, b = ctx.saved_tensors
        g1, g2 = torch.empty_like(a), torch.empty_like(b)
        ctx.be.call('be_etas2depth_bwd', a, b, _f32c(gz, a.device), a.numel(), g1, g2)
        return g1, g2, None


class Inverse3(Function):
    """:104-112"""

    @staticmethod
    def forward(ctx, A, be):
        a = _f32c(A, A.device)
        inv = torch.empty_like(a)
        be.call('be_inverse_3by3', a, a.numel() // 9, inv)
        ctx.save_for_backward(inv)
        ctx.be = be
        return inv

    @staticmethod
    def backward(ctx, g):
        (inv,) = ctx.saved_tensors
        gA = torch.empty_like(inv)
        ctx.be.call('be_inverse_3by3_bwd', inv, _f32c(g, inv.device), inv.numel() // 9, gA)
        return gA, None


class ImageDerivative(Function):
    """:114-117.  [N,3,H,W] -> [N,3,H-2,W-2]"""

    @staticmethod
    def forward(ctx, img, be):
        x = _f32c(img, img.device)
        H, W = x.shape[-2], x.shape[-1]
        out = torch.empty(tuple(x.shape[:-2]) + (H - 2, W - 2), device=x.device, dtype=torch.float32)
        be.call('be_image_derivative', x, x.numel() // (H * W), H, W, out)
        ctx.save_for_backward(x)
        ctx.be = be
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        H, W = x.shape[-2], x.shape[-1]
        gx = torch.empty_like(x)
        ctx.be.call('be_image_derivative_bwd', x, _f32c(g, x.device), x.numel() // (H * W), H, W, gx)
        return gx, None


class Fold(Function):
    """Overlap sum of patches [n,R,R,Hp,Wp] -> [n,H,W], divided by num_patches (mode 0); backward = the matching Unfold."""

    @staticmethod
    def forward(ctx, patches, be, n, H, W, mode):
        p = _f32c(patches, patches.device)
        out = torch.empty(n, H, W, device=p.device, dtype=torch.float32)
        be.call('be_fold', p, n, mode, out)
        ctx.be, ctx.n, ctx.mode, ctx.shape = be, n, mode, tuple(patches.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        gp = torch.empty(ctx.shape, device=g.device, dtype=torch.float32)
        ctx.be.call('be_unfold', _f32c(g, g.device), ctx.n, ctx.mode, gp)
        return gp, None, None, None, None, None
