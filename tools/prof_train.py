#!/usr/bin/env python
"""Small driver for profiling the training step (GlobalLossFused fwd+bwd) under ncu: `python tools/prof_train.py [pairs] [steps]`."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch  # noqa: E402
import synth  # noqa: E402
from blurry_edges_b200 import GlobalLossFused  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
same = len(sys.argv) > 3 and sys.argv[3] == 'same'      # the training call of the reference: criteria(est, img_gt, img_gt, ...)
S, L = 147, 4096
cam = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], mag=4.0, rho_prime=10.39, cam_params=cam,
                          batch_size=B, gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                          gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                          gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])
crit = GlobalLossFused(args, None, 'cuda:0')
crit.update_gamma()
raw = synth.raw_global(B, L, seed=300).cuda().requires_grad_(True)
img = synth.image_pairs(B, S, S, seed=301).cuda()
gt, bd, deri, zg = [t.cuda() for t in synth.loss_targets(B, S, S, seed=302)]
if same:
    img = gt
for _ in range(3):
    raw.grad = None
    crit(raw, img, gt, bd, deri, zg).backward()
torch.cuda.synchronize()
crit.ctx.set_timing(True)
t0 = time.perf_counter()
for _ in range(steps):
    raw.grad = None
    crit(raw, img, gt, bd, deri, zg).backward()
torch.cuda.synchronize()
print('kernel ms (memset, setup, run3<TRAINFWD>, normalise, pack, loss2, reduce):', ' '.join(f'{m:.4f}' for m in crit.ctx.last_train_timing()))
print(f'B={B}{" same-gt" if same else ""}: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms/step wall, {B * L * steps / (time.perf_counter() - t0) / 1e6:.1f} M patches/s')
