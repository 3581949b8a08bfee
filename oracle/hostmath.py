"""ctypes wrapper of oracle/_build/libbe_hostmath.so (TEST INFRASTRUCTURE: the kernels' fp32 arithmetic run on the
host, multi-threaded).  Used by tests/ and by bench.py's CPU-baseline legs only."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import be_oracle as O
from .build_oracle import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _cam7(cam: O.Camera, rho_prime):
    return np.array([cam.numerator, cam.k_fac, cam.k_const, cam.k_root, cam.intercept, cam.s, rho_prime], dtype=np.float32)


def _f(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def colors(est, img_planar, g: O.Geometry, cam: O.Camera, param_mode=2):
    """est [M,L,10], img [M,3,H,W] -> [M,3,3,Hp,Wp]"""
    est, img = _f(est), _f(img_planar)
    M = est.shape[0]
    strides = np.array([3 * g.H * g.W, 0, g.H * g.W, g.W, 1], dtype=np.int64)
    out = np.empty((M, 3, 3, g.Hp, g.Wp), dtype=np.float32)
    rc = lib().behm_colors(_p(est), C.c_int(param_mode), _p(img), _p(strides), M, g.H, g.W, g.R, g.stride, C.c_float(g.w),
                           C.c_float(g.lam), _p(_cam7(cam, 10.39)), _p(out))
    assert rc == 0
    return out


def render_fold(est, img_planar, g: O.Geometry, cam: O.Camera, rho_prime=10.39, densify=None, param_mode=0):
    """est [B,L,12], img [B,2,3,H,W] -> six maps (numpy fp32)"""
    est, img = _f(est), _f(img_planar)
    B, H, W = est.shape[0], g.H, g.W
    strides = np.array([6 * H * W, 3 * H * W, H * W, W, 1], dtype=np.int64)
    outs = [np.empty(s, dtype=np.float32) for s in ((B, 2, 3, H, W), (B, 3, H, W), (B, 3, H, W), (B, 1, H, W), (B, H, W), (B, H, W))]
    rc = lib().behm_render_fold(_p(est), C.c_int(param_mode), _p(img), _p(strides), B, H, W, g.R, g.stride, C.c_float(g.w),
                                C.c_float(g.lam), _p(_cam7(cam, rho_prime)), C.c_int(int(densify == 'w')), *[_p(o) for o in outs])
    assert rc == 0
    return outs


def global_loss(raw, img_ny, img_gt, bndry_dist, deri, bndry_depth, gammas, g: O.Geometry, cam: O.Camera):
    """fp32 host run of the loss kernel's algorithm: returns (loss, terms[7], grad [B,L,12], gimg [B,2,3,H,W], gbnd [B,H,W])."""
    raw, img_ny, img_gt, bd, deri, zg = map(_f, (raw, img_ny, img_gt, bndry_dist, deri, bndry_depth))
    B, H, W = raw.shape[0], g.H, g.W
    gam = np.asarray(gammas, dtype=np.float32)
    terms = np.zeros(7, np.float32)
    loss = np.zeros(1, np.float32)
    grad = np.zeros_like(raw)
    gimg = np.zeros((B, 2, 3, H, W), np.float32)
    gbnd = np.zeros((B, H, W), np.float32)
    rc = lib().behm_global_loss(_p(raw), _p(img_ny), _p(img_gt), _p(bd), _p(deri), _p(zg), _p(gam), B, H, W, g.R, g.stride,
                                C.c_float(g.w), C.c_float(g.lam), _p(_cam7(cam, 10.39)), 0, _p(terms), _p(loss), _p(grad),
                                _p(gimg), _p(gbnd))
    assert rc == 0
    return float(loss[0]), terms, grad, gimg, gbnd


def local_loss(est, img_ny, gt_img, bndry_dist, deri, betas, g: O.Geometry, cam: O.Camera):
    """returns (loss, terms[3], grad [B,10])"""
    est, img_ny, gt_img, bd, deri = map(_f, (est, img_ny, gt_img, bndry_dist, deri))
    B, R = est.shape[0], g.R
    gam = np.asarray([1.0, betas[0], betas[1], 0, 0, 0, 0], dtype=np.float32)
    terms = np.zeros(7, np.float32)
    loss = np.zeros(1, np.float32)
    grad = np.zeros_like(est)
    dummy = np.zeros(1, np.float32)
    rc = lib().behm_global_loss(_p(est), _p(img_ny), _p(gt_img), _p(bd), _p(deri), _p(dummy), _p(gam), B, R, R, R, g.stride,
                                C.c_float(g.w), C.c_float(g.lam), _p(_cam7(cam, 10.39)), 1, _p(terms), _p(loss), _p(grad),
                                _p(dummy), _p(dummy))
    assert rc == 0
    return float(loss[0]), terms[:3], grad


def pack_selfcheck(params12, cam: O.Camera, xy, w=1.0, rho_prime=10.39):
    """Packed two-pixel functions (be_pack.cuh) vs the scalar specification (be_math.cuh) on the pixel positions xy [n,2]
    of one patch with restored parameters params12 -> 5 error figures (see hm_pack_selfcheck)."""
    p, xy = _f(params12), _f(xy)
    out = np.zeros(5, dtype=np.float32)
    lib().hm_pack_selfcheck(_p(p), _p(_cam7(cam, rho_prime)), _p(xy), C.c_int(xy.shape[0]), C.c_float(w), _p(out))
    return out
