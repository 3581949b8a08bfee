import sys, argparse
sys.path[:0] = ['/root/repo', '/root/repo/tests']
import torch, synth, bench
from blurry_edges_b200 import Context, _lib, make_config, GlobalLossFused
S, B = 147, 64
est, img = bench.make_inputs(B, seed=5)
est, img = est.cuda(), img.cuda()
ctx = Context(make_config(H=S, W=S, max_batch=B), 'cuda:0')
lay = _lib.planar_layout(S, S)
first = [o.clone() for o in ctx.render_fold(est, img, lay)]
worst = 0.0
for it in range(400):
    out = ctx.render_fold(est, img, lay)
    if it % 20 == 19:
        for a, b in zip(first, out):
            worst = max(worst, float((a - b).abs().max() / a.abs().max()))
torch.cuda.synchronize()
print('inference soak: 400 launches, worst rel diff vs first', worst)
cam = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], mag=4.0, rho_prime=10.39, cam_params=cam,
                          batch_size=32, gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                          gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                          gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])
crit = GlobalLossFused(args, None, 'cuda:0'); crit.update_gamma()
raw = synth.raw_global(32, 4096, seed=300).cuda().requires_grad_(True)
im = synth.image_pairs(32, S, S, seed=301).cuda()
gt, bd, deri, zg = [t.cuda() for t in synth.loss_targets(32, S, S, seed=302)]
raw.grad = None; l0 = crit(raw, im, gt, bd, deri, zg); l0.backward(); g0 = raw.grad.clone(); l0 = l0.item()
wl = wg = 0.0
for it in range(150):
    raw.grad = None
    l = crit(raw, im, gt, bd, deri, zg); l.backward()
    if it % 10 == 9:
        wl = max(wl, abs(l.item() - l0) / abs(l0)); wg = max(wg, float((raw.grad - g0).abs().max() / g0.abs().max()))
print('training soak: 150 steps, worst rel loss diff', wl, 'worst grad diff', wg, 'finite', bool(torch.isfinite(raw.grad).all()))
