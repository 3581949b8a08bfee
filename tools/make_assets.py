#!/usr/bin/env python
"""Synthetic weights and datasets with the file names and layouts the reference's entry scripts load: the
pre-trained weights and the datasets of the reference are not in its repository, so the scripts are exercised with random-init
state_dicts (torch.manual_seed(0)) and with the basic-shape scenes of tests/golden/shapes147.npz (reference generator) /
hash-synthesised arrays of tests/synth.py.

  python tools/make_assets.py <reference dir> <out dir>

  <out>/weights/pretrained_{local_stage,global_stage,global_stage_w}.pth            blurry_edges_test*.py:183-190, global_data_pre_cal.py:64
  <out>/eval/{images_ny,alphas,depth_maps}.npy          2 pairs 147x147             data/dataset.py:58-73
  <out>/big/{images_ny,alphas,depth_maps}.npy           1 pair 235x235 (2x2 blocks)
  <out>/train/*_{train,val}.npy                          4 + 2 scenes               data/dataset.py:19-36 (mode 'global' / 'global_pre')
  <out>/train/patches/*_{train,val}.npy                  128 + 64 patches 21x21     data/dataset.py:10-18 (mode 'local')"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)
import synth  # noqa: E402


def main():
    ref, out = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2])
    sys.path.insert(0, ref)
    import models
    w = os.path.join(out, 'weights')
    for d in (w, os.path.join(out, 'eval'), os.path.join(out, 'big'), os.path.join(out, 'train', 'patches')):
        os.makedirs(d, exist_ok=True)
    torch.manual_seed(0)
    torch.save(models.LocalStage().state_dict(), os.path.join(w, 'pretrained_local_stage.pth'))
    g = models.GlobalStage(in_parameter_size=38, out_parameter_size=12, device='cpu').state_dict()
    torch.save(g, os.path.join(w, 'pretrained_global_stage.pth'))
    torch.save(g, os.path.join(w, 'pretrained_global_stage_w.pth'))

    a = synth.shapes_arrays()
    S = a['clean'].shape[2]
    # ---- evaluation sets ----
    np.save(os.path.join(out, 'eval', 'images_ny.npy'), a['noisy'][:2])
    np.save(os.path.join(out, 'eval', 'alphas.npy'), a['alpha'][:2])
    np.save(os.path.join(out, 'eval', 'depth_maps.npy'), 0.75 + 0.43 * synth.u01((2, S, S), 52))
    np.save(os.path.join(out, 'big', 'images_ny.npy'), synth.photon_pairs(1, 235, 235, seed=61))
    np.save(os.path.join(out, 'big', 'alphas.npy'), np.array([190.0]))
    np.save(os.path.join(out, 'big', 'depth_maps.npy'), 0.75 + 0.43 * synth.u01((1, 235, 235), 62))
    # ---- global-stage training set (train_val_data_generator.py:137-185 file set + global_data_pre_cal.py:33) ----
    L = ((S - 21) // 2 + 1) ** 2
    deri = np.zeros(a['clean'].shape)
    deri[:, :, 1:-1, 1:-1, :] = a['deri']
    gt = a['clean'] / 255 * a['alpha'][:, None, None, None, None]
    for part, sl in (('train', slice(0, 4)), ('val', slice(4, 6))):
        n = sl.stop - sl.start
        t = os.path.join(out, 'train')
        np.save(f'{t}/images_ny_{part}.npy', a['noisy'][sl])
        np.save(f'{t}/images_gt_{part}.npy', gt[sl])
        np.save(f'{t}/alphas_{part}.npy', a['alpha'][sl])
        np.save(f'{t}/derivative_maps_{part}.npy', deri[sl])
        np.save(f'{t}/boundary_distances_{part}.npy', a['dist'][sl])
        np.save(f'{t}/boundary_depths_{part}.npy', a['depth'][sl])
        np.save(f'{t}/params_src_{part}.npy', synth.normalish((n, 2, L, 19), 90 + sl.start, 0.1, torch.float64).numpy())
    # ---- local-stage training set: 21x21 crops of the scenes (train_val_data_generator.py:187-275 file set) ----
    for part, n, seed in (('train', 128, 95), ('val', 64, 96)):
        pos = (synth.u01((n, 3), seed) * [a['clean'].shape[0], S - 21, S - 21]).astype(int)
        crop = lambda arr, k, m=None: arr[pos[k, 0]][..., pos[k, 1]:pos[k, 1] + 21, pos[k, 2]:pos[k, 2] + 21, :] if arr.ndim == 5 else \
            arr[pos[k, 0], pos[k, 1]:pos[k, 1] + 21, pos[k, 2]:pos[k, 2] + 21]
        t = os.path.join(out, 'train', 'patches')
        np.save(f'{t}/patches_ny_{part}.npy', np.stack([crop(a['noisy'], k)[k % 2] for k in range(n)]))
        np.save(f'{t}/patches_gt_{part}.npy', np.stack([crop(gt, k)[k % 2] for k in range(n)]))
        np.save(f'{t}/alphas_{part}.npy', a['alpha'][pos[:, 0]])
        np.save(f'{t}/boundary_distances_{part}.npy', np.stack([crop(a['dist'], k) for k in range(n)]))
        np.save(f'{t}/derivative_maps_{part}.npy', np.stack([crop(deri, k)[k % 2] for k in range(n)]))
    print(f'assets written under {out}')


if __name__ == '__main__':
    main()
