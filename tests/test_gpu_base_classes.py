"""The reference's helper-class METHODS on this library's kernels (blurry_edges_b200.base / ops): every method forward
and backward against the oracle (fp64 + autograd), and a script-level subclass composed exactly like
blurry_edges_test.PostProcess (restated here) producing the six maps through the method-granularity path."""
import argparse

import numpy as np
import pytest
import torch

import synth
from common import F32, F64, GEOMS, MAPS, geom, inference_inputs, relmax
from oracle import be_oracle as O

pytestmark = pytest.mark.gpu
CAM = O.Camera()
CAMP = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}


def _args(S, B=1, **kw):
    return argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], batch_size=B, mag=4.0, rho_prime=10.39,
                              densify=None, cam_params=CAMP, **kw)


def _both(fn_ours, fn_ref, inputs, tol_f=1e-5, tol_g=5e-5, seed=0):
    """forward + backward (random upstream gradient) of ours (cuda fp32) vs the oracle (cpu fp64)"""
    xs = [t.clone().cuda().requires_grad_(True) for t in inputs]
    xr = [t.clone().to(F64).requires_grad_(True) for t in inputs]
    yo, yr = fn_ours(*xs), fn_ref(*xr)
    assert tuple(yo.shape) == tuple(yr.shape)
    assert relmax(yo.detach().cpu().numpy(), yr.detach().numpy()) < tol_f
    up = synth.uniform(tuple(yr.shape), 77 + seed, -1.0, 1.0)
    go = torch.autograd.grad(yo, xs, up.cuda())
    gr = torch.autograd.grad(yr, xr, up.to(F64))
    for a, b in zip(go, gr):
        assert relmax(a.cpu().numpy(), b.numpy()) < tol_g


@pytest.fixture(scope='module')
def gbase():
    from blurry_edges_b200 import PostProcessGlobalBase

    class G(PostProcessGlobalBase):
        pass
    return G(_args(45), 'cuda:0')


@pytest.fixture(scope='module')
def lbase():
    from blurry_edges_b200 import PostProcessLocalBase

    class Lc(PostProcessLocalBase):
        pass
    return Lc(_args(45, B=6), 'cuda:0')


def _geo_global(B, Hp, seed):
    p = O.restore_global(synth.raw_global(B, Hp * Hp, seed=seed))[..., :8]          # [B,L,8]
    return p.permute(0, 2, 1).reshape(B, 8, Hp, Hp).contiguous()


def test_attributes_match_reference_definitions(gbase):
    g = geom(45)
    assert (gbase.H_patches, gbase.W_patches) == (g.Hp, g.Wp)
    assert torch.equal(gbase.num_patches.cpu(), O.cover_count(g))
    assert gbase.x.shape == (1, 21, 21, 1, 1) and gbase.ridge.shape == (1, 1, 1, 3, 3)
    assert float(gbase.ridge[0, 0, 0, 0, 0]) == pytest.approx(g.lam)
    assert torch.equal(gbase.sobel_x[0, 0].cpu(), torch.tensor([[-1., 0, 1], [-2, 0, 2], [-1, 0, 1]]))
    assert torch.equal(gbase.sobel_y[0, 0].cpu(), torch.tensor([[1., 2, 1], [0, 0, 0], [-1, -2, -1]]))


def test_params2dists_global_and_local(gbase, lbase):
    g = geom(45)
    ref = lambda p: O.wedge_distances(p.permute(0, 2, 3, 1).reshape(-1, 8), 21).reshape(p.shape[0], g.Hp, g.Wp, 2, 21, 21).permute(0, 3, 4, 5, 1, 2)
    _both(gbase.params2dists, ref, [_geo_global(2, g.Hp, 5)])
    pl = O.restore_global(synth.raw_global(1, 6, seed=9))[0, :, :8].contiguous()                        # [6,8]
    _both(lbase.params2dists, lambda p: O.wedge_distances(p, 21), [pl])


def test_params2dists_accepts_channel_slices(gbase):
    g = geom(45)
    full = torch.cat([_geo_global(1, g.Hp, 5), synth.uniform((1, 4, g.Hp, g.Wp), 3)], 1).cuda()        # [1,12,Hp,Wp] like est in the scripts
    a = gbase.params2dists(full[:, :8, :, :])
    b = gbase.params2dists(full[:, :8, :, :].contiguous())
    assert torch.equal(a, b)


def test_dists2indicators_and_etas(gbase, lbase):
    g = geom(45)
    d = synth.uniform((2, 2, 21, 21, g.Hp, g.Wp), 11, -1.5, 1.5)
    e = synth.uniform((2, 2, g.Hp, g.Wp), 12, 0.02, 0.5)
    ref = lambda dd, ee: O.soft_indicators(dd.permute(0, 4, 5, 1, 2, 3).reshape(-1, 2, 21, 21), ee.permute(0, 2, 3, 1).reshape(-1, 2)) \
        .reshape(2, g.Hp, g.Wp, 3, 21, 21).permute(0, 3, 4, 5, 1, 2)
    _both(gbase.dists2indicators, ref, [d, e])
    _both(lbase.dists2indicators, O.soft_indicators, [synth.uniform((6, 2, 21, 21), 13, -1.5, 1.5), synth.uniform((6, 2), 14, 0.02, 0.5)])
    _both(gbase.params2etas, O.eta_from_coef, [synth.uniform((2, 4, g.Hp, g.Wp), 15, -0.5, 1.0)])
    _both(lambda x: gbase.normalized_gaussian(x), O.bump, [synth.uniform((2, 21, 21, 5, 5), 16, -0.3, 0.3)])
    _both(lambda x: gbase.normalized_gaussian(x, 0.2), lambda x: O.bump(x, 0.2), [synth.uniform((3, 7), 17, -0.5, 0.5)])


def test_inverse_sobel_depth(gbase):
    from blurry_edges_b200 import DepthEtas
    A = synth.uniform((2, 5, 5, 3, 3), 21, -1.0, 1.0)
    A = A @ A.transpose(-1, -2) + 4.86 * torch.eye(3)
    _both(gbase.inverse_3by3, O.inv3_sym, [A])
    # get_adjA(A, A2, trA, trA2) (utils/postprocessing_loss.py:127-128,148-149): adj(A) = det(A) A^-1
    Ad = A.double()
    adj = gbase.get_adjA(A.cuda(), None, None, None).cpu().double()
    ref = torch.linalg.det(Ad)[..., None, None] * torch.linalg.inv(Ad)
    assert float((adj - ref).abs().max() / ref.abs().max()) < 1e-5
    _both(gbase.get_image_derivative, O.sobel_mag, [synth.uniform((4, 3, 21, 21), 22)])
    _both(gbase.get_image_derivative, O.sobel_mag, [synth.uniform((2, 3, 45, 45), 23)])
    cal = DepthEtas(_args(45), 'cuda:0')
    assert float(cal.intercept) == pytest.approx(CAM.intercept, rel=2e-7) and cal.numerator == pytest.approx(CAM.numerator)
    e1, e2 = synth.uniform((2, 13, 13), 24, 0.01, 0.6), synth.uniform((2, 13, 13), 25, 0.01, 0.6)
    _both(cal.etas2depth, lambda a, b: O.depth_from_etas(CAM, a, b), [e1, e2])
    _both(lambda z: cal.depth2sigma(z, 10.39), lambda z: O.refocus_sigma(CAM, z, 10.39), [synth.uniform((2, 13, 13), 26, 0.7, 1.2)])


def test_folds(gbase):
    g = geom(45)
    B = 1
    pat = synth.uniform((B, 2, 3, 21, 21, g.Hp, g.Wp), 31)
    n = O.cover_count(g, F64)
    to_pm = lambda t, C: t.reshape(-1, C, 21, 21, g.L).permute(0, 4, 1, 2, 3).reshape(-1, C, 21, 21)   # [M*L,C,R,R]
    ref_color = lambda p: (O.fold_sum(to_pm(p.reshape(B * 2, 3, 21, 21, g.Hp, g.Wp), 3), B * 2, g) / n).reshape(B, 2, 3, g.H, g.W)
    _both(gbase.local2global_color, ref_color, [pat])
    ref_b = lambda p: O.fold_sum(to_pm(p, 1), B, g) / n
    _both(gbase.local2global_bndry, ref_b, [synth.uniform((B, 1, 21, 21, g.Hp, g.Wp), 32)])
    dm = synth.uniform((B, 21, 21, g.Hp, g.Wp), 33, 0.7, 1.2)
    mk = (synth.uniform((B, 21, 21, g.Hp, g.Wp), 34) * 3).floor().to(torch.int32)
    depth, conf = gbase.local2global_depth(dm.cuda(), mk.cuda())
    cnt = O.fold_sum(to_pm((mk > 0).to(F64).unsqueeze(1), 1), B, g)[:, 0]
    dref = O.fold_sum(to_pm(dm.to(F64).unsqueeze(1), 1), B, g)[:, 0] / torch.where(cnt > 0, cnt, torch.ones_like(cnt))
    assert relmax(depth.cpu().numpy(), dref.numpy()) < 1e-5 and relmax(conf.cpu().numpy(), (cnt / n).numpy()) < 1e-6


def test_script_level_subclass_composed_from_methods():
    """A PostProcess written against the base-class methods, composed as blurry_edges_test.py:19-100 composes them."""
    from blurry_edges_b200 import DepthEtas, PostProcessGlobalBase

    class PostProcess(PostProcessGlobalBase):
        def __init__(self, args, cal, device):
            super().__init__(args, device)
            self.cal, self.rho_prime = cal, args.rho_prime

        def colors(self, wedges, pix):                       # normal equations over both images, 3 wedges x 3 channels
            A = wedges.permute(0, 5, 6, 1, 3, 4, 2).reshape(self.batch_size, self.H_patches, self.W_patches, -1, 3)
            y = pix.permute(0, 5, 6, 1, 3, 4, 2).reshape(self.batch_size, self.H_patches, self.W_patches, -1, 3)
            At = A.transpose(-1, -2)
            return (self.inverse_3by3(At @ A + self.ridge) @ (At @ y)).permute(0, 4, 3, 1, 2)

        def forward(self, est, pix):
            est = est.permute(0, 2, 1).view(self.batch_size, 12, self.H_patches, self.W_patches)
            etas = self.params2etas(est[:, 8:])
            dists = self.params2dists(est[:, :8])
            w1, w2 = self.dists2indicators(dists, etas[:, :2]), self.dists2indicators(dists, etas[:, 2:])
            col = self.colors(torch.stack([w1, w2], 1), pix)
            paint = lambda w: (w.unsqueeze(1) * col.unsqueeze(-3).unsqueeze(-3)).sum(2)
            z1, z2 = self.cal.etas2depth(etas[:, 0], etas[:, 2]), self.cal.etas2depth(etas[:, 1], etas[:, 3])
            m = (self.normalized_gaussian(dists[:, 0]) > 0.5).to(torch.int32)
            t = (self.normalized_gaussian(dists[:, 1]) > 0.5).to(torch.int32) * 2
            m = torch.where((t == 2) | (dists[:, 1] >= 0), t, m)
            dmap = torch.where(m == 1, z1[:, None, None], torch.where(m == 2, z2[:, None, None], torch.zeros_like(dists[:, 0])))
            a1, a2 = dists[:, 0].abs(), dists[:, 1].abs()
            lb = self.normalized_gaussian(torch.where(dists[:, 1] >= 0, dists[:, 1], torch.where(a1 < a2, a1, a2)))
            sharp = paint(self.dists2indicators(dists, torch.full_like(etas[:, :2], 1e-4)))
            s1 = torch.where((m == 1).sum((1, 2)) > 0, self.cal.depth2sigma(z1, self.rho_prime), torch.full_like(z1, 1e-4))
            s2 = torch.where((m == 2).sum((1, 2)) > 0, self.cal.depth2sigma(z2, self.rho_prime), torch.full_like(z2, 1e-4))
            refoc = paint(self.dists2indicators(dists, torch.stack([s1, s2], 1)))
            depth, conf = self.local2global_depth(dmap, m)
            return (self.local2global_color(torch.stack([paint(w1), paint(w2)], 1)), self.local2global_color(sharp, pair=False),
                    self.local2global_color(refoc, pair=False), self.local2global_bndry(lb.unsqueeze(1)), depth, conf)

    S = GEOMS['mid']
    g, est, img = inference_inputs('mid', 'normal', F32)
    args = _args(S)
    helper = PostProcess(args, DepthEtas(args, 'cuda:0'), 'cuda:0')
    pix = torch.nn.Unfold(21, stride=2)(img[0]).view(1, 2, 3, 21, 21, g.Hp, g.Wp).cuda()
    with torch.no_grad():
        maps = helper(est.cuda(), pix)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, None)
    tol = {'image': 2e-5, 'sharp': 5e-5, 'refoc': 2e-5, 'bndry': 1e-5, 'depth': 1e-5, 'conf': 1e-6}    # bmm in fp32 on top of the kernels
    for name, r, o in zip(MAPS, ref, maps):
        assert relmax(o.cpu().numpy(), r.numpy()) < tol[name], name


def test_edge_frame_helpers_match_the_reference_formulas(gbase, lbase):
    """dist4edge / dist4axial / itemize_params (utils/postprocessing_loss.py:26-41): d = -sin a (X - x) + cos a (Y - y),
    t = cos a (X - x) + sin a (Y - y) on the pixel grid, in both broadcasting layouts."""
    g = geom(45)
    p = _geo_global(2, g.Hp, 5).cuda()
    x0, y0, _, _, a0, _, _, _ = gbase.itemize_params(p)
    ax = O.pixel_axis(21, F64)
    X, Y = ax.view(1, 1, 21, 1, 1), ax.view(1, 21, 1, 1, 1)
    xs, ys, an = [t.cpu().to(F64) for t in (x0, y0, a0)]
    d_ref = -torch.sin(an) * (X - xs) + torch.cos(an) * (Y - ys)
    t_ref = torch.cos(an) * (X - xs) + torch.sin(an) * (Y - ys)
    assert relmax(gbase.dist4edge(x0, y0, a0).cpu().numpy(), d_ref.numpy()) < 1e-6
    assert relmax(gbase.dist4axial(x0, y0, a0).cpu().numpy(), t_ref.numpy()) < 1e-6
    pl = O.restore_global(synth.raw_global(1, 6, seed=9))[0, :, :8].contiguous().cuda()
    xl, yl, _, _, al, _, _, _ = lbase.itemize_params(pl)
    dl = lbase.dist4edge(xl, yl, al)
    assert dl.shape == (6, 21, 21)
    ref = -torch.sin(al.cpu().to(F64)) * (ax.view(1, 1, 21) - xl.cpu().to(F64)) + torch.cos(al.cpu().to(F64)) * (ax.view(1, 21, 1) - yl.cpu().to(F64))
    assert relmax(dl.cpu().numpy(), ref.numpy()) < 1e-6
