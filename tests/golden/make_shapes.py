"""Config-3 inputs as SURVEY.md 8(d) prescribes them: basic-shape scenes from the reference's OWN generator
(train_val_data_generator.SyntheticShapeDataGenerator, imported read-only from /root/reference, never copied), with its
noise model, plus golden loss / gradient of the unmodified GlobalLoss on them.

Run in the build container only:   python tests/golden/make_shapes.py
Writes tests/golden/shapes147.npz:
  clean  uint8  [N,2,147,147,3]   `imgs` of generate_synthetic_image (:31-116; rounded to integers there, :110)
  noisy  uint8  [N,2,147,147,3]   clip(round(Poisson(clean/255*alpha) + 2 N(0,1)), 0, alpha)              (:165-185)
  dist   uint16 [N,147,147]       boundary_dist (city-block distance to the nearest drawn boundary, :98-108)
  depth  float64[N,147,147]       boundary_depth (:74-84)
  alpha  float64[N]
and the goldens (first NG scenes, raw = tests/synth.raw_global(NG, L, seed=81), gamma_idx 0):
  train/{loss,grad}   criteria(est, img_gt, img_gt, ...)   the training call   (global_training.py:210), fp64
  val/{loss,grad}     criteria(est, img_ny, img_gt, ...)   the validation call (global_training.py:166), fp64
  train32/{loss,grad} the training call run by the unmodified reference in fp32 (its own rounding noise, for the record)
The derivative maps are NOT stored: they are a deterministic function of `clean` (:111-115) that tests/synth.shapes_batch
recomputes exactly (integer Sobel sums, one sqrt, one division)."""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refimport  # noqa: E402
import synth  # noqa: E402

N, NG, SEED = 8, 2, 1869


def main():
    assert refimport.available(), 'reference not mounted'
    utils = refimport.module('utils')
    gen_mod = refimport.module('train_val_data_generator')
    tmp = tempfile.mkdtemp()
    args = refimport.get_args('data_gen_train_val', ['--data_path', tmp])
    utils.set_seed(SEED)
    gen = gen_mod.SyntheticShapeDataGenerator(args)
    num_obj = np.random.randint(args.num_shape[0], args.num_shape[1], size=N)                 # :137
    clean, dist, depth, deri = [], [], [], []
    for n in num_obj:
        imgs, _aif, _loc, _zimg, bdepth, bdist, d = gen.generate_synthetic_image(int(n))          # :31-116
        clean.append(imgs); dist.append(bdist); depth.append(bdepth); deri.append(d)
    clean, dist, depth, deri = np.stack(clean), np.stack(dist), np.stack(depth), np.stack(deri)
    alpha = np.random.rand(N) * (args.alpha[1] - args.alpha[0]) + args.alpha[0]                   # :171
    noisy = np.zeros_like(clean)
    for i in range(N):                                                                            # :172-179
        for ii in range(2):
            prime = clean[i, ii] / 255 * alpha[i]
            ny = np.random.poisson(prime).astype(float) + args.sigma * np.random.randn(*prime.shape)
            noisy[i, ii] = ny.clip(0, alpha[i]).round()
    assert clean.min() >= 0 and clean.max() <= 255 and np.array_equal(clean, clean.round())
    assert noisy.min() >= 0 and noisy.max() <= 255 and np.array_equal(noisy, noisy.round())
    assert dist.min() >= 0 and np.array_equal(dist, dist.round())
    out = {'clean': clean.astype(np.uint8), 'noisy': noisy.astype(np.uint8), 'dist': dist.astype(np.uint16), 'depth': depth, 'alpha': alpha}
    # the loader's recomputed derivative must be the generator's own, bit for bit
    np.savez_compressed(os.path.join(HERE, 'shapes147.npz'), **out)
    chk = synth.shapes_arrays()
    assert np.array_equal(chk['deri'], deri[:, :, 1:-1, 1:-1, :]), 'derivative restatement differs from the generator'

    # goldens: the unmodified GlobalLoss on the first NG scenes (data/dataset.py:27-36,50-56 conversions inside shapes_batch)
    gt_mod = refimport.module('global_training')
    targs = refimport.get_args('global_train', ['--cuda', 'cpu', '--batch_size', NG])
    for dt, tag in ((torch.float64, ''), (torch.float32, '32')):
        cal = utils.DepthEtas(targs, 'cpu')
        crit = refimport.to_dtype(gt_mod.GlobalLoss(targs, cal, 'cpu'), cal, dt)
        crit.update_gamma()
        L = crit.H_patches * crit.W_patches
        ny, gt, bd, dv, zg = synth.shapes_batch(NG, dtype=dt)
        for name, first in (('train', gt), ('val', ny)):
            if tag and name != 'train':
                continue
            raw = synth.raw_global(NG, L, seed=81, dtype=dt).requires_grad_(True)
            loss = crit(raw, first, gt, bd, dv, zg)                                              # global_training.py:147-157
            (grad,) = torch.autograd.grad(loss, raw)
            out[f'{name}{tag}.loss'] = loss.detach().numpy()
            out[f'{name}{tag}.grad'] = grad.numpy()
            print(name + tag, float(loss))
        out['gammas'] = np.array([getattr(crit, a) for a in ('gamma_color', 'gamma_color_cons', 'gamma_bndry_cons', 'gamma_smthns',
                                                              'gamma_smthns_cons', 'gamma_bndry_loc', 'gamma_depth')])
    path = os.path.join(HERE, 'shapes147.npz')
    np.savez_compressed(path, **out)
    print(f'{path}: {os.path.getsize(path) / 1024:.0f} KiB')


if __name__ == '__main__':
    main()
