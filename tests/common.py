"""Shared helpers for the test-suite: golden loader, regenerated inputs, error measures."""
from __future__ import annotations

import math
import os

import numpy as np
import torch

import synth
from oracle import be_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GEOMS = {'tiny': 29, 'mid': 45}
F64, F32 = torch.float64, torch.float32
MAPS = ('image', 'sharp', 'refoc', 'bndry', 'depth', 'conf')


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, f'{name}.npz'))

    def __call__(self, key):
        return self.z[key.replace('/', '.')]

    def has(self, key):
        return key.replace('/', '.') in self.z.files


def geom(S):
    return O.Geometry(H=S, W=S)


def planar_pair(img_b2hw3):
    """[B,2,H,W,3] -> [B,2,3,H,W]"""
    return img_b2hw3.permute(0, 1, 4, 2, 3).contiguous()


def relmax(a, b):
    """max|a-b| / max|b| (the 'relative to tensor max' measure of SURVEY section 4)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def inference_inputs(gname, kind, dt):
    S = GEOMS[gname]
    g = geom(S)
    img = planar_pair(synth.image_pairs(1, S, S, seed=3, dtype=dt))
    # restored in fp32 as the reference's script does (blurry_edges_test.py:135-138), then cast: every dtype sees the same values
    est = O.restore_global(synth.raw_global(1, g.L, seed=7, kind=kind, dtype=F32)).to(dt)
    return g, est, img


def gloss_inputs(gname, kind, dt, B=2):
    S = GEOMS[gname]
    g = geom(S)
    img_ny = synth.image_pairs(B, S, S, seed=31, dtype=dt)
    img_gt, bd, deri, zgt = synth.loss_targets(B, S, S, seed=31, dtype=dt)
    raw = synth.raw_global(B, g.L, seed=33, kind=kind, dtype=dt)
    return g, raw, img_ny, img_gt, bd, deri, zgt


def shapes_inference_inputs(dt):
    """Inputs of tests/golden/shapes147_infer.npz (made by tests/golden/make_shapes_inference.py): the noisy pair of one full-size scene
    of the reference's generator, est restored in fp32 as the script does, est10 for pass A."""
    gold = Golden('shapes147_infer')
    g = geom(147)
    img = planar_pair(synth.shapes_batch(1, first=int(gold('scene')), dtype=dt)[0])                # [1,2,3,H,W]
    est = O.restore_global(synth.raw_global(1, g.L, seed=83, dtype=F32)).to(dt)
    est10 = synth.est_local(2, g.L, seed=85, dtype=dt)
    return gold, g, est, est10, img
