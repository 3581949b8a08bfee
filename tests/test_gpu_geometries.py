"""Geometries other than the reference defaults (R=21, stride=2, square 147x147, w=1): the kernels take R <= 21, any stride < R,
rectangular images and any cap width w (utils/args.py:9-15 are only defaults).  These run the generic code paths - column
residues per warp, run splitting, threads without pixels, the run-time-R variant of the loss kernel - against the fp64 oracle with
the same tolerances as the default-geometry tests."""
import argparse

import numpy as np
import pytest
import torch

import synth
from common import F32, F64, MAPS, planar_pair, relmax
from oracle import be_oracle as O

pytestmark = pytest.mark.gpu
CAMP = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
TOL = dict(image=1e-5, sharp=2e-5, refoc=1e-5, bndry=1e-5, depth=1e-5, conf=1e-5)
GEOS = {
    'R11_s3_rect': dict(R=11, stride=3, H=35, W=41, w=1.0),
    'R21_s1_rect': dict(R=21, stride=1, H=27, W=33, w=1.0),
    'R7_s2': dict(R=7, stride=2, H=15, W=15, w=1.0),
    'R21_s5_w05': dict(R=21, stride=5, H=41, W=46, w=0.5),
    'R15_s4_w2': dict(R=15, stride=4, H=31, W=39, w=2.0),
    # even patch sizes (no centre pixel, R*R even: the two pixel slots of the last thread are both in use or both empty)
    'R12_s2_rect': dict(R=12, stride=2, H=22, W=26, w=1.0),
    'R20_s3_rect': dict(R=20, stride=3, H=35, W=41, w=1.0),
    # runs of 1, 2 and 3 patches: the prologue / drain of the depth-2 pipeline and its three hand-off buffers
    'one_patch': dict(R=21, stride=2, H=21, W=21, w=1.0),
    'two_patches': dict(R=21, stride=2, H=21, W=23, w=1.0),
    'three_by_three': dict(R=21, stride=2, H=25, W=25, w=1.0),
}


def _geo(name):
    k = GEOS[name]
    return O.Geometry(R=k['R'], stride=k['stride'], H=k['H'], W=k['W'], w=k['w']), O.Camera(R=k['R'])


@pytest.mark.parametrize('name', list(GEOS))
@pytest.mark.parametrize('densify', [None, 'w'])
def test_pass_b_other_geometries(name, densify):
    from blurry_edges_b200 import Context, _lib, make_config
    g, cam = _geo(name)
    B = 2
    est = O.restore_global(synth.raw_global(B, g.L, seed=51))
    img = planar_pair(synth.image_pairs(B, g.H, g.W, seed=52))
    ctx = Context(make_config(R=g.R, stride=g.stride, H=g.H, W=g.W, w=g.w, max_batch=B), 'cuda:0')
    out = ctx.render_fold(est.cuda(), img.cuda(), _lib.planar_layout(g.H, g.W), densify_w=(densify == 'w'))
    ref = O.inference(est.to(F64), img.to(F64), g, cam, 10.39, densify)
    for n, r, o in zip(MAPS, ref, out):
        if n == 'depth' or n == 'conf':
            # the discrete mask may flip where |d| is within fp32 rounding of a threshold: compare where confidence agrees
            same = (o.cpu().double() - r).abs() <= TOL[n] * float(r.abs().max())
            assert float(same.double().mean()) > 0.999, n
        else:
            assert relmax(o.cpu().numpy(), r.numpy()) < TOL[n], n


@pytest.mark.parametrize('name', list(GEOS))
def test_pass_a_other_geometries(name):
    from blurry_edges_b200 import Context, _lib, make_config
    g, cam = _geo(name)
    M = 3
    raw = synth.raw_global(M, g.L, seed=53)
    est10 = O.restore_global(raw)[..., :10].contiguous()
    img = synth.image_pairs(M, g.H, g.W, seed=54)[:, 0].permute(0, 3, 1, 2).contiguous()       # [M,3,H,W]
    ctx = Context(make_config(R=g.R, stride=g.stride, H=g.H, W=g.W, w=g.w, max_batch=M), 'cuda:0')
    col = ctx.colors(est10.cuda(), img.cuda(), _lib.single_planar_layout(g.H, g.W), _lib.PARAMS_LOCAL10)
    ref = O.colors_only(est10.to(F64), img.to(F64), g)
    assert relmax(col.cpu().numpy(), ref.numpy()) < 1e-5


@pytest.mark.parametrize('name', list(GEOS))
def test_global_loss_other_geometries(name):
    from blurry_edges_b200 import GlobalLossFused
    g, cam = _geo(name)
    B = 2
    ranges = dict(gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                  gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                  gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])
    args = argparse.Namespace(R=g.R, stride=g.stride, w=g.w, alpha_lambda=5e-3, img_size=[g.H, g.W], batch_size=B, mag=4.0,
                              cam_params=CAMP, **ranges)
    crit = GlobalLossFused(args, None, 'cuda:0')
    crit.update_gamma()
    gam = crit.gammas()
    img_ny = synth.image_pairs(B, g.H, g.W, seed=61)
    img_gt, bd, deri, zgt = synth.loss_targets(B, g.H, g.W, seed=61)
    raw = synth.raw_global(B, g.L, seed=63)
    est = raw.clone().cuda().requires_grad_(True)
    loss = crit(est, img_ny.cuda(), img_gt.cuda(), bd.cuda(), deri.cuda(), zgt.cuda())
    loss.backward()
    r64 = raw.to(F64).requires_grad_(True)
    l64 = O.global_loss(r64, img_ny.to(F64), img_gt.to(F64), bd.to(F64), deri.to(F64), zgt.to(F64), gam, g, cam)
    (g64,) = torch.autograd.grad(l64, r64)
    assert abs(loss.item() - l64.item()) <= 5e-6 * abs(l64.item())
    got, ref = est.grad.cpu().double().numpy(), g64.numpy()
    emax = float(np.abs(got - ref).max() / np.abs(ref).max())
    el2 = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    assert emax < 5e-5 and el2 < 2e-5, (emax, el2)


def test_pixels_no_patch_covers_behave_as_in_the_reference():
    """(H - R) not divisible by the stride leaves a border no patch covers: nn.Fold(ones) is 0 there, so the reference's averaged maps
    are 0/0 = NaN and its depth (divided by max(count, 1)) is 0 (utils/postprocessing_loss.py:151-173).  Same pattern here, and the
    covered pixels are unaffected."""
    from blurry_edges_b200 import Context, _lib, make_config
    g, cam = O.Geometry(R=21, stride=2, H=24, W=26, w=1.0), O.Camera(R=21)     # row 23 and column 25 are uncovered
    est = O.restore_global(synth.raw_global(1, g.L, seed=55))
    img = planar_pair(synth.image_pairs(1, g.H, g.W, seed=56))
    ctx = Context(make_config(R=21, stride=2, H=24, W=26, w=1.0, max_batch=1), 'cuda:0')
    out = ctx.render_fold(est.cuda(), img.cuda(), _lib.planar_layout(g.H, g.W))
    ref = O.inference(est.to(F64), img.to(F64), g, cam, 10.39, None)
    for n, r, o in zip(MAPS, ref, out):
        o = o.cpu().double()
        assert torch.equal(torch.isnan(o), torch.isnan(r)), n
        ok = ~torch.isnan(r)
        if n == 'depth':
            assert float(o[..., 23, :].abs().max()) == 0.0 and float(o[..., :, 25].abs().max()) == 0.0
        else:
            assert bool(torch.isnan(r[..., 23, :]).all()) and bool(torch.isnan(r[..., :, 25]).all()), n
        assert float((o[ok] - r[ok]).abs().max()) <= 2e-5 * float(r[ok].abs().max()), n
