"""Golden INFERENCE outputs of the unmodified reference on a full-size scene of its own generator (hard edges, flat regions, its
Poisson + Gaussian noise model): blurry_edges_test.PostProcess (imported read-only from /root/reference, never copied) on the noisy
pair of scene 2 of tests/golden/shapes147.npz, in fp64.

Run in the build container only, after make_shapes.py:   python tests/golden/make_shapes_inference.py
Writes tests/golden/shapes147_infer.npz:
  scene                       index of the scene in shapes147.npz
  passA                       PostProcess(est10, patches, colors_only=True)     [2,3,3,64,64]     (blurry_edges_test.py:81-92)
  {image,sharp,bndry}         averaged maps of colors_only=False that do not depend on the mask rule   (blurry_edges_test.py:93-100)
  none/{refoc,depth,conf}, w/{refoc,depth,conf}   refocused image (its sigma falls back where a wedge owns no mask pixel, :66-72),
                              depth and confidence under the default rule and --densify w              (:47-57,95-99)
Inputs are not stored: the image pair comes from shapes147.npz (tests/synth.shapes_batch), est = restore(raw_global(1, L, seed=83))
restored in fp32 as the script does (:135-138), est10 = est_local(2, L, seed=85)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refimport  # noqa: E402
import synth  # noqa: E402
from make_golden import restore, unfold_pair  # noqa: E402

SCENE, S = 2, 147
F64 = torch.float64


def main():
    assert refimport.available(), 'reference not mounted'
    bt = refimport.module('blurry_edges_test')
    utils = refimport.module('utils')
    out = {'scene': np.array(SCENE)}
    ny = synth.shapes_batch(1, first=SCENE, dtype=F64)[0]                       # [1,2,H,W,3], divided by alpha (data/dataset.py:50)
    for densify in (None, 'w'):
        argv = ['--cuda', 'cpu', '--img_size', S, S, '--batch_size', 1] + (['--densify', densify] if densify else [])
        args = refimport.get_args('eval', argv)
        cal = utils.DepthEtas(args, 'cpu')
        pp = refimport.to_dtype(bt.PostProcess(args, cal, 'cpu'), cal, F64)
        L = pp.H_patches * pp.W_patches
        pat = unfold_pair(ny, args.R, args.stride)
        est = restore(synth.raw_global(1, L, seed=83, dtype=torch.float32)).to(F64)
        maps = pp(est, pat, colors_only=False)
        names = ('image', 'sharp', 'refoc', 'bndry', 'depth', 'conf')
        for n, m in zip(names, maps):
            if n in ('refoc', 'depth', 'conf'):
                out[f'{densify or "none"}/{n}'] = np.asarray(m)
            elif densify is None:
                out[n] = np.asarray(m)
            else:
                assert np.array_equal(out[n], np.asarray(m)), n                  # the averaged maps do not depend on the rule
        if densify is None:
            estA = synth.est_local(2, L, seed=85, dtype=F64)
            out['passA'] = pp(estA, pat, colors_only=True).detach().numpy()
    path = os.path.join(HERE, 'shapes147_infer.npz')
    np.savez_compressed(path, **{k.replace('/', '.'): v for k, v in out.items()})       # tests/common.Golden maps '/' to '.'
    print(f'{path}: {os.path.getsize(path) / 1024:.0f} KiB', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
