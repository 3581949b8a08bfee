#!/usr/bin/env python
"""Tiny end-to-end exercise of every hot-path kernel, meant for compute-sanitizer where it is available
   (compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_smoke.py); it is closed on the round-1 GPU pool,
   so tests/test_gpu_*::test_*repeatability* are the race checks that actually ran."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import torch  # noqa: E402
import synth  # noqa: E402
from blurry_edges_b200 import Context, GlobalLossFused, LocalLossFused, _lib, make_config, smish  # noqa: E402

cam = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
for (R, stride, H, W) in ((21, 2, 29, 33), (11, 3, 26, 23)):
    B = 2
    Hp, Wp = (H - R) // stride + 1, (W - R) // stride + 1
    L = Hp * Wp
    ctx = Context(make_config(R=R, stride=stride, H=H, W=W, max_batch=B), 'cuda:0')
    raw = synth.raw_global(B, L, seed=1).cuda()
    img = synth.image_pairs(B, H, W, seed=2).cuda()
    planar = img.permute(0, 1, 4, 2, 3).contiguous()
    out = ctx.render_fold(raw, planar, _lib.planar_layout(H, W), param_mode=_lib.PARAMS_RAW12)
    col = ctx.colors(raw[..., :10].contiguous().repeat(2, 1, 1), planar.view(2 * B, 3, H, W), _lib.single_planar_layout(H, W), _lib.PARAMS_LOCALRAW10)
    host = ctx.host_render_fold(raw.cpu().pin_memory(), planar.cpu().pin_memory(), _lib.planar_layout(H, W), param_mode=_lib.PARAMS_RAW12)
    args = argparse.Namespace(R=R, stride=stride, w=1.0, alpha_lambda=5e-3, img_size=[H, W], mag=4.0, rho_prime=10.39, cam_params=cam,
                              batch_size=B, gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                              gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                              gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])
    crit = GlobalLossFused(args, None, 'cuda:0')
    crit.update_gamma()
    gt, bd, deri, zg = [t.cuda() for t in synth.loss_targets(B, H, W, seed=3)]
    est = raw.clone().requires_grad_(True)
    crit(est, img, gt, bd, deri, zg).backward()
    torch.cuda.synchronize()
    print(f'R={R} stride={stride} {H}x{W}: maps {float(out[0].abs().sum()):.4f} colours {float(col.abs().sum()):.4f} grad {float(est.grad.abs().sum()):.6f}')
    # round-2 paths: fixed-order fold (slabs + stage reduce), host-buffer training entry (two kernel streams), single-launch local loss
    ctx.set_deterministic(True)
    out_d = ctx.render_fold(raw, planar, _lib.planar_layout(H, W), param_mode=_lib.PARAMS_RAW12)
    ctx.set_deterministic(False)
    crit.deterministic = True
    est2 = raw.clone().requires_grad_(True)
    crit(est2, gt, gt, bd, deri, zg).backward()
    crit.deterministic = False
    ht = crit.ctx.host_global_loss(raw.cpu(), gt.cpu(), gt.cpu(), bd.cpu(), deri.cpu(), zg.cpu(), crit.gammas())
    torch.cuda.synchronize()
    print(f'  deterministic maps {float(out_d[0].abs().sum()):.4f} grad {float(est2.grad.abs().sum()):.6f} host loss {float(ht[1]):.6f}')
largs = argparse.Namespace(R=21, w=1.0, alpha_lambda=5e-3, batch_size=8, mag=4.0, cam_params=cam, beta_bndry_loc=0.001, beta_smthns=0.0005,
                           dynamic_epoch=200)
lcrit = LocalLossFused(largs, 'cuda:0')
lcrit.final_beta()
le, lny, lgt, lbd, lderi = [t.cuda() for t in synth.local_batch(8, 21, seed=41)]
leaf = le.clone().requires_grad_(True)
for _ in range(2):                                   # twice: the ticket of the last-CTA reduction must have reset itself
    leaf.grad = None
    lcrit(leaf * 1.0, lny, lgt, lbd, lderi).backward()
torch.cuda.synchronize()
print(f'local loss grad {float(leaf.grad.abs().sum()):.6f} terms {lcrit.terms.tolist()}')
x = torch.linspace(-5, 5, 1001, device='cuda', requires_grad=True)
smish(x).sum().backward()
torch.cuda.synchronize()
print('smish ok')
