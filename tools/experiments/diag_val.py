"""Diagnostic: validation-call loss of the reference GlobalLoss (fp32 / fp64, on cuda:0) vs GlobalLossFused, with est from a
xavier-initialised GlobalStage as global_training.py produces it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, 'baseline', '_ref', 'Blurry-Edges')
sys.path[:0] = [os.path.join(ROOT, 'tests', '_stubs'), REF, ROOT, os.path.join(ROOT, 'tests')]
import numpy as np, torch
import synth
sys.argv = ['x', '--cuda', 'cuda:0', '--batch_size', '2']
import utils, models, global_training as gtm
args = utils.get_args('global_train')
dev = torch.device('cuda:0')
utils.set_seed(1898)
net = models.GlobalStage(in_parameter_size=args.input_size, out_parameter_size=args.output_size, device=dev).to(dev)
for p in net.parameters():
    if p.dim() > 1:
        torch.nn.init.xavier_normal_(p)
net.eval()
B = 2
L = 4096
param = synth.normalish((B, 2, L, 19), 94, 0.1).to(dev)
ny, gt, bd, deri, zg = [t.to(dev) for t in synth.shapes_batch(B, first=4)]
with torch.no_grad():
    est = net(param.permute(0, 2, 1, 3).flatten(2, 3))
print('est stats: min/max/std per group', [(float(est[..., a:b].min()), float(est[..., a:b].max()), float(est[..., a:b].std())) for a, b in ((0, 4), (4, 8), (8, 12))])
def ref_loss(dt, final):
    cal = utils.DepthEtas(args, dev)
    crit = gtm.GlobalLoss(args, cal, dev)
    for k in ('x', 'y', 'ridge', 'num_patches', 'sobel_x', 'sobel_y'):
        setattr(crit, k, getattr(crit, k).to(dt))
    for k in ('intercept', 'theta_mid', 'theta_wng'):
        setattr(cal, k, getattr(cal, k).to(dt))
    crit.update_gamma()
    if final: crit.final_gamma()
    with torch.no_grad():
        l = crit(est.to(dt), ny.to(dt), gt.to(dt), bd.to(dt), deri.to(dt), zg.to(dt))
        # the individual terms
        terms = None
    return float(l)
from blurry_edges_b200 import GlobalLossFused
for final in (False, True):
    ours = GlobalLossFused(args, None, dev)
    ours.update_gamma()
    if final: ours.final_gamma()
    with torch.no_grad():
        lo = float(ours(est, ny, gt, bd, deri, zg))
    l32, l64 = ref_loss(torch.float32, final), ref_loss(torch.float64, final)
    print(f'final={final}: ref64 {l64:.10f} ref32 {l32:.10f} (rel {abs(l32-l64)/l64:.2e})  ours {lo:.10f} (rel {abs(lo-l64)/l64:.2e}) terms {ours.terms.tolist()} gammas {ours.gammas()}')

# gradients of the training call (clean image twice, gamma_idx 0) w.r.t. est
def ref_grad(dt):
    cal = utils.DepthEtas(args, dev)
    crit = gtm.GlobalLoss(args, cal, dev)
    for k in ('x', 'y', 'ridge', 'num_patches', 'sobel_x', 'sobel_y'):
        setattr(crit, k, getattr(crit, k).to(dt))
    for k in ('intercept', 'theta_mid', 'theta_wng'):
        setattr(cal, k, getattr(cal, k).to(dt))
    crit.update_gamma()
    e = est.to(dt).clone().requires_grad_(True)
    l = crit(e, gt.to(dt), gt.to(dt), bd.to(dt), deri.to(dt), zg.to(dt))
    (g,) = torch.autograd.grad(l, e)
    return g.double()
g64, g32 = ref_grad(torch.float64), ref_grad(torch.float32)
ours = GlobalLossFused(args, None, dev)
ours.update_gamma()
e = est.clone().requires_grad_(True)
ours(e, gt, gt, bd, deri, zg).backward()
go = e.grad.double()
def err(a, b):
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())
print('grad: ref32 vs ref64 (max rel-to-max, rel-L2):', err(g32, g64), ' ours vs ref64:', err(go, g64), ' nan in ref64/ref32/ours:', int(torch.isnan(g64).sum()), int(torch.isnan(g32).sum()), int(torch.isnan(go).sum()))
# clip-normalised direction the optimiser sees: cosine between implementations
cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()))
print('cosine ref32/ref64', cos(g32, g64), ' ours/ref64', cos(go, g64))
