"""Platform-independent synthetic inputs (integer hash -> exact fp32 values), so that golden
outputs can be committed without their inputs and regenerated bit-identically anywhere."""
from __future__ import annotations

import math

import numpy as np
import torch


def u01(shape, seed: int) -> np.ndarray:
    """Uniform-looking values k/2^24 in [0,1), float64 array (each exactly representable in fp32)."""
    n = int(np.prod(shape))
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(0x9E3779B1) + np.uint64(0x7F4A7C15)) & np.uint64(0xFFFFFFFF)
    for mul in (0x45D9F3B, 0x119DE1F3):
        h ^= h >> np.uint64(16)
        h = (h * np.uint64(mul)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return ((h & np.uint64(0xFFFFFF)).astype(np.float64) / float(1 << 24)).reshape(shape)


def uniform(shape, seed, lo=0.0, hi=1.0, dtype=torch.float32) -> torch.Tensor:
    return torch.from_numpy(lo + (hi - lo) * u01(shape, seed)).to(torch.float32).to(dtype)


def normalish(shape, seed, scale=1.0, dtype=torch.float32) -> torch.Tensor:
    """Irwin-Hall(4) stand-in for N(0,1): unit variance, support +-3.46."""
    s = sum(u01(shape, seed * 4 + k) for k in range(4))
    return torch.from_numpy((s - 2.0) * math.sqrt(3.0) * scale).to(torch.float32).to(dtype)


# ---- input families (all fp32-exact, returned in `dtype`) -------------------------------
def raw_global(B, L, seed, kind='normal', dtype=torch.float32):
    """Raw GlobalStage-like output [B,L,12]: 0.1*N(0,1) ('normal', realistic) or U[-1,1] ('stress')."""
    if kind == 'normal':
        return normalish((B, L, 12), seed, 0.1, dtype)
    return uniform((B, L, 12), seed, -1.0, 1.0, dtype)


def est_local(M, L, seed, dtype=torch.float32):
    """Local-stage style params [M,L,10]: xy in [-1.2,1.2], angles in [0,2pi), eta coef in [-0.5,1]."""
    xy = uniform((M, L, 4), seed, -1.2, 1.2)
    ang = uniform((M, L, 4), seed + 1, 0.0, 6.28125)
    eta = uniform((M, L, 2), seed + 2, -0.5, 1.0)
    return torch.cat([xy, ang, eta], -1).to(dtype)


def image_pairs(B, H, W, seed, dtype=torch.float32):
    """[B,2,H,W,3] in [0,1): smooth-ish blobs + hash noise so that colours are non-trivial."""
    base = u01((B, 2, H, W, 3), seed)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    blob = 0.5 + 0.5 * np.sin(0.11 * xx[None, None, :, :, None] + 0.07 * yy[None, None, :, :, None]
                              + np.arange(3)[None, None, None, None, :] + np.arange(B)[:, None, None, None, None])
    img = np.round((0.7 * blob + 0.3 * base) * 4096) / 4096
    return torch.from_numpy(img).to(torch.float32).to(dtype)


def photon_pairs(N, H, W, seed, alpha=190):
    """uint8-exact photon-count style pairs [N,2,H,W,3] (float64 counts in 0..alpha), like images_ny.npy."""
    return np.floor(image_pairs(N, H, W, seed, torch.float64).numpy() * alpha).clip(0, alpha)


def loss_targets(B, H, W, seed, dtype=torch.float32):
    """(img_gt [B,2,H,W,3], bndry_dist [B,H,W], deri [B,2,H-2,W-2,3], bndry_depth [B,H,W])."""
    img_gt = image_pairs(B, H, W, seed + 10, dtype)
    bd = uniform((B, H, W), seed + 11, 0.0, 6.0, dtype)
    deri = uniform((B, 2, H - 2, W - 2, 3), seed + 12, 0.0, 2.0, dtype)
    z = uniform((B, H, W), seed + 13, 0.75, 1.18)
    keep = uniform((B, H, W), seed + 14) < 0.4
    return img_gt, bd, deri, torch.where(keep, z, torch.zeros_like(z)).to(dtype)


def local_batch(B, R, seed, dtype=torch.float32):
    """LocalLoss inputs: est [B,10] raw, img_ny/gt [B,R,R,3], bndry_dist [B,R,R], deri [B,R-2,R-2,3]."""
    est = torch.cat([uniform((B, 4), seed, -1.2, 1.2), uniform((B, 4), seed + 1, -7.0, 7.0),
                     uniform((B, 2), seed + 2, -0.5, 1.0)], 1).to(dtype)
    ny = image_pairs(B, R, R, seed + 3, dtype)[:, 0]
    gt = image_pairs(B, R, R, seed + 4, dtype)[:, 1]
    bd = uniform((B, R, R), seed + 5, 0.0, 6.0, dtype)
    deri = uniform((B, R - 2, R - 2, 3), seed + 6, 0.0, 2.0, dtype)
    return est, ny, gt, bd, deri


# ---- basic-shape scenes of the reference's own generator (tests/golden/shapes147.npz, made by tests/golden/make_shapes.py) ----
_SHAPES = None


def shapes_arrays():
    """The committed fixture as float64 arrays + the derivative maps recomputed from the clean images exactly as the generator
    defines them (train_val_data_generator.py:111-115: sqrt(Sobel_x^2 + Sobel_y^2) / 255 per channel; interior pixels only, as
    data/dataset.py:32 keeps them).  The Sobel sums are integers, so the result does not depend on the summation order."""
    global _SHAPES
    if _SHAPES is None:
        import os
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'shapes147.npz'))
        c = z['clean'].astype(np.float64)                                   # [N,2,H,W,3]
        gx = (c[:, :, :-2, 2:] - c[:, :, :-2, :-2]) + 2 * (c[:, :, 1:-1, 2:] - c[:, :, 1:-1, :-2]) + (c[:, :, 2:, 2:] - c[:, :, 2:, :-2])
        gy = (c[:, :, :-2, :-2] + 2 * c[:, :, :-2, 1:-1] + c[:, :, :-2, 2:]) - (c[:, :, 2:, :-2] + 2 * c[:, :, 2:, 1:-1] + c[:, :, 2:, 2:])
        _SHAPES = {'clean': c, 'noisy': z['noisy'].astype(np.float64), 'dist': z['dist'].astype(np.float64), 'depth': z['depth'],
                   'alpha': z['alpha'], 'deri': np.sqrt(gx ** 2 + gy ** 2) / 255, 'z': z}
    return _SHAPES


def shapes_batch(B, first=0, dtype=torch.float32):
    """(img_ny, img_gt [B,2,H,W,3], bndry_dist [B,H,W], deri [B,2,H-2,W-2,3], bndry_depth [B,H,W]) as ShapeDataset(mode='global')
    hands them to the training loop (data/dataset.py:27-36,50-56: float32 arrays, images divided by alpha); scenes are taken
    round-robin from the fixture starting at `first`."""
    a = shapes_arrays()
    idx = (first + np.arange(B)) % a['clean'].shape[0]
    f32 = lambda x: torch.from_numpy(x[idx]).float()
    alpha = f32(a['alpha']).view(B, 1, 1, 1, 1)
    img_gt = f32(a['clean'] / 255 * a['alpha'][:, None, None, None, None]) / alpha      # images_gt = clean/255*alpha (:175), then /alpha
    img_ny = f32(a['noisy']) / alpha
    return tuple(t.to(dtype) for t in (img_ny, img_gt, f32(a['dist']), f32(a['deri']), f32(a['depth'])))
