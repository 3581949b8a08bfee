"""What a non-default patch size costs: the loss kernel has a compile-time R = 21 variant (neighbour offsets
of the stencil are immediates) and a run-time-R variant for every other R <= 21; R > 21 is refused (two pixel slots per render thread
cover at most 448 pixels).  Times the global-loss training step and inference pass B per patch and per patch pixel on 147x147 pairs.

  python tools/experiments/geometry_cost.py [pairs]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch

import synth
from blurry_edges_b200 import Context, GlobalLossFused, _lib, make_config

CAMP = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
RANGES = dict(gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
              gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
              gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = 'cuda:0'
    H = W = 147
    print(f'{B} pairs of {H}x{W}; ms per call, ns per patch, ps per patch pixel')
    for R, s in ((21, 2), (19, 2), (17, 2), (15, 2), (11, 2), (21, 4), (21, 1)):
        Hp, Wp = (H - R) // s + 1, (W - R) // s + 1
        L = Hp * Wp
        args = argparse.Namespace(R=R, stride=s, w=1.0, alpha_lambda=5e-3, img_size=[H, W], batch_size=B, mag=4.0, cam_params=CAMP, **RANGES)
        crit = GlobalLossFused(args, None, dev)
        crit.update_gamma()
        gt, bd, deri, zg = [t.to(dev) for t in synth.loss_targets(B, H, W, seed=61)]
        raw = synth.raw_global(B, L, seed=63).to(dev)

        def step():
            est = raw.clone().requires_grad_(True)
            crit(est, gt, gt, bd, deri, zg).backward()

        t_train = timed(step)
        ctx = Context(make_config(R=R, stride=s, H=H, W=W, max_batch=B), dev)
        from oracle import be_oracle as O      # parameter restore only (tools/, not the product path)
        est = O.restore_global(raw.cpu()).to(dev)
        img = synth.image_pairs(B, H, W, seed=52).permute(0, 1, 4, 2, 3).contiguous().to(dev)
        lay = _lib.planar_layout(H, W)
        t_inf = timed(lambda: ctx.render_fold(est, img, lay))
        n = B * L
        print(f'R={R:2d} stride={s}: {L:6d} patches/pair | train {t_train:7.3f} ms {t_train * 1e6 / n:7.2f} ns/patch {t_train * 1e9 / (n * R * R):6.1f} ps/px'
              f' | pass B {t_inf:7.3f} ms {t_inf * 1e6 / n:7.2f} ns/patch {t_inf * 1e9 / (n * R * R):6.1f} ps/px')
        del crit, ctx


if __name__ == '__main__':
    main()
