"""GPU parity tests of pass A / pass B (call through the C ABI via blurry_edges_b200).

Tolerance (north_star: "within 1e-5 relative in fp32"): max|ours - oracle_fp64| / max|oracle_fp64| <= 1e-5 per map.
The eta=1e-4 "sharpened" render turns a 1-ulp distance rounding into an O(1e-3) change of one wedge value at the
(measure-zero) pixels within ~1e-6 of an edge, so that map gets 2e-5 after the fold average; the reference's own
fp32 path is 1e-5..1.4e-3 away from its fp64 path on the same inputs (tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

import synth
from common import F32, F64, GEOMS, MAPS, Golden, geom, inference_inputs, planar_pair, relmax
from oracle import be_oracle as O

pytestmark = pytest.mark.gpu
CAM = O.Camera()
TOL = {'image': 1e-5, 'sharp': 2e-5, 'refoc': 1e-5, 'bndry': 1e-5, 'depth': 1e-5, 'conf': 1e-6}


def _ctx(S, max_batch=1):
    from blurry_edges_b200 import Context, make_config
    return Context(make_config(H=S, W=S, max_batch=max_batch), 'cuda:0')


def _run_b(ctx, est, img, densify=None):
    from blurry_edges_b200 import _lib
    S = img.shape[-1]
    out = ctx.render_fold(est.cuda().contiguous(), img.cuda().contiguous(), _lib.planar_layout(S, S), densify_w=(densify == 'w'))
    torch.cuda.synchronize()
    return [o.cpu() for o in out]


@pytest.mark.parametrize('gname', list(GEOMS))
def test_pass_a_vs_oracle_and_golden(gname):
    from blurry_edges_b200 import _lib
    S = GEOMS[gname]
    g = geom(S)
    img = planar_pair(synth.image_pairs(1, S, S, seed=3))[0]
    est = synth.est_local(2, g.L, seed=5)
    ctx = _ctx(S)
    got = ctx.colors(est.cuda(), img.cuda().contiguous(), _lib.single_planar_layout(S, S)).cpu().numpy()
    ref = O.colors_only(est.to(F64), img.to(F64), g).numpy()
    assert relmax(got, ref) < 1e-5
    assert relmax(got, Golden('inference')(f'{gname}/passA/f64')) < 1e-5      # the unmodified reference, fp64


@pytest.mark.parametrize('gname', list(GEOMS))
@pytest.mark.parametrize('densify', [None, 'w'])
@pytest.mark.parametrize('kind', ['normal', 'stress'])
def test_pass_b_vs_oracle_and_golden(gname, densify, kind):
    g, est, img = inference_inputs(gname, kind, F32)
    got = _run_b(_ctx(GEOMS[gname]), est, img, densify)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, densify)
    gold = Golden('inference')
    for name, r, o in zip(MAPS, ref, got[:6]):
        assert relmax(o.numpy(), r.numpy()) < TOL[name], name
        # the unmodified reference run in fp64 on the same (fp32-restored) parameters: same bound as against the oracle
        assert relmax(o.numpy(), gold(f'{gname}/passB/{densify or "none"}/{kind}/f64/{name}')) < TOL[name], name
    thres = 0.0 if densify == 'w' else 0.05
    thr = torch.where(ref[5] > thres, ref[4], torch.zeros_like(ref[4]))
    close = (ref[5] - thres).abs() < 1e-6                                     # confidence within rounding of the threshold
    assert (((got[6].double() - thr).abs() < 1e-5) | close).all()


@pytest.mark.parametrize('densify', [None, 'w'])
def test_generator_scene_147_vs_oracle_and_reference_golden(densify):
    """Full-size scene of the reference's own generator (hard edges, flat regions, Poisson + Gaussian noise; tests/golden/shapes147.npz):
    pass A and pass B against the fp64 oracle and against the unmodified reference's fp64 outputs (shapes147_infer.npz), same bounds.
    Depth / confidence: the discrete mask may flip where |d| is within fp32 rounding of a threshold, which moves one patch's vote at a
    pixel - compared where the vote pattern agrees, which must be all but a handful of pixels."""
    from blurry_edges_b200 import _lib
    from common import shapes_inference_inputs
    gold, g, est, est10, img = shapes_inference_inputs(F32)
    ctx = _ctx(147)
    if densify is None:
        col = ctx.colors(est10.cuda(), img[0].cuda().contiguous(), _lib.single_planar_layout(147, 147)).cpu().numpy()
        assert relmax(col, O.colors_only(est10.to(F64), img[0].to(F64), g).numpy()) < 1e-5
        assert relmax(col, gold('passA')) < 1e-5
    got = _run_b(ctx, est, img, densify)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, densify)
    for name, r, o in zip(MAPS, ref, got[:6]):
        gd = gold(f'{densify or "none"}/{name}' if name in ('refoc', 'depth', 'conf') else name)
        for target in (r.numpy(), gd):
            if name in ('depth', 'conf'):
                same = np.abs(o.numpy().astype(np.float64) - target) <= TOL['depth'] * np.abs(target).max()
                assert same.mean() > 0.9995, (name, 1 - same.mean())
            else:
                assert relmax(o.numpy(), target) < TOL[name], name


@pytest.mark.parametrize('densify', [None, 'w'])
def test_config1_147_maps_and_depth_metrics(densify):
    """Config 1: same `est` the unchanged reference driver produced; maps vs the fp64 oracle, depth metrics
    (delta1..3, RMSE, AbsRel of utils/metrics.py) equal to the reference's printed values to 4 decimals."""
    from test_oracle_golden import _eval_depth
    gc = Golden('config1')
    S = 147
    g = geom(S)
    img = planar_pair(torch.from_numpy(synth.photon_pairs(1, S, S, seed=51, alpha=190)).float() / 190.0)
    est = torch.from_numpy(gc('config1/est'))
    got = _run_b(_ctx(S), est, img, densify)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, densify)
    for name, r, o in zip(MAPS, ref, got[:6]):
        assert relmax(o.numpy(), r.numpy()) < TOL[name], name
    gt = synth.uniform((1, S, S), 52, 0.75, 1.18).numpy().astype(np.float64)
    depth = got[6].numpy().astype(np.float64)
    ours = _eval_depth(depth, gt, depth > 0)
    np.testing.assert_allclose(ours, gc(f'config1/{densify or "none"}/metrics'), rtol=0, atol=5e-5)


def test_layouts_and_unfolded_input_agree():
    """planar, dataset-native channels-last and the reference's unfolded patches give the same maps."""
    from blurry_edges_b200 import _lib
    S = GEOMS['mid']
    g, est, img = inference_inputs('mid', 'normal', F32)
    ctx = _ctx(S)
    a = _run_b(ctx, est, img)
    cl = img.permute(0, 1, 3, 4, 2).contiguous().cuda()                      # [B,2,H,W,3]
    b = ctx.render_fold(est.cuda(), cl, _lib.channels_last_layout(S, S))
    unf = torch.nn.Unfold(g.R, stride=g.stride)(img[0]).view(2, 3, g.R, g.R, g.Hp, g.Wp).cuda().contiguous()
    refolded = ctx.refold(unf)
    assert torch.equal(refolded.cpu(), img[0])
    c = ctx.render_fold(est.cuda(), refolded, _lib.planar_layout(S, S))
    for x, y, z in zip(a, b, c):
        assert relmax(y.cpu().numpy(), x.numpy()) < 2e-6 and relmax(z.cpu().numpy(), x.numpy()) < 2e-6


def test_cover_count_matches_fold_of_ones():
    for S in (29, 45, 147):
        assert torch.equal(_ctx(S).cover_count().cpu(), O.cover_count(geom(S)))


def test_fused_class_mirrors_reference_call_sites():
    """PostProcessFused takes exactly what blurry_edges_test.py:128,140 pass (unfolded patches) and returns
    colours [2,3,3,Hp,Wp] / six NumPy maps."""
    import argparse
    from blurry_edges_b200 import PostProcessFused
    S = GEOMS['mid']
    g, est, img = inference_inputs('mid', 'normal', F32)
    args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], batch_size=1, mag=4.0, rho_prime=10.39,
                              densify=None, cam_params={'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6})
    helper = PostProcessFused(args, None, 'cuda:0')
    pat = torch.nn.Unfold(g.R, stride=g.stride)(img[0]).view(2, 3, g.R, g.R, g.Hp, g.Wp).cuda()
    estA = synth.est_local(2, g.L, seed=5)
    colors = helper(estA.cuda(), pat, colors_only=True)
    assert colors.shape == (2, 3, 3, g.Hp, g.Wp)
    assert relmax(colors.cpu().numpy(), O.colors_only(estA.to(F64), img[0].to(F64), g).numpy()) < 1e-5
    maps = helper(est.cuda(), pat, colors_only=False)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, None)
    assert all(isinstance(m, np.ndarray) for m in maps)
    for name, r, o in zip(MAPS, ref, maps):
        assert o.shape == tuple(r.shape) and relmax(o, r.numpy()) < TOL[name], name


def test_host_buffer_entry_point_equals_device_entry_point():
    from blurry_edges_b200 import _lib
    S = GEOMS['mid']
    g, est, img = inference_inputs('mid', 'normal', F32)
    B = 3
    est = torch.cat([est, est.flip(1), est * 0.5]).contiguous()
    img = torch.cat([img, img.flip(-1), 1 - img]).contiguous()
    ctx = _ctx(S, max_batch=B)
    dev = _run_b(ctx, est, img)
    host = ctx.host_render_fold(est.pin_memory(), img.pin_memory(), _lib.planar_layout(S, S))
    for d, h in zip(dev, host):
        assert relmax(h.numpy(), d.numpy()) < 2e-6
    six = ctx.host_render_fold(est.pin_memory(), img.pin_memory(), _lib.planar_layout(S, S), want_thresholded=False)
    assert len(six) == 6 and all(relmax(a.numpy(), b.numpy()) < 2e-6 for a, b in zip(six, host))
    # pageable output arrays (synchronous staging inside cudaMemcpyAsync) give the same result
    pageable = [torch.full_like(h, float('nan')).clone() for h in host]
    assert not any(t.is_pinned() for t in pageable)
    ctx.host_render_fold(est, img, _lib.planar_layout(S, S), out=pageable)
    for p, h in zip(pageable, host):
        assert relmax(p.numpy(), h.numpy()) < 2e-6      # the fold's reduction order differs from call to call


@pytest.mark.parametrize('B', [8, 11])
def test_host_buffer_entry_point_chunked_pipeline(B):
    """More pairs than pipeline chunks (8): exercises both kernel streams, the 16-byte (B=8) and the scalar (B=11, odd
    chunk sizes) forms of the export kernel, and repeated calls into the same pinned outputs."""
    from blurry_edges_b200 import _lib
    S = GEOMS['tiny']
    g = geom(S)
    est = O.restore_global(synth.raw_global(B, g.L, seed=91))
    img = planar_pair(synth.image_pairs(B, S, S, seed=92))
    ctx = _ctx(S, max_batch=B)
    dev = _run_b(ctx, est, img)
    host = ctx.host_render_fold(est.pin_memory(), img.pin_memory(), _lib.planar_layout(S, S))
    host = ctx.host_render_fold(est.pin_memory(), img.pin_memory(), _lib.planar_layout(S, S), out=host)
    for d, h in zip(dev, host):
        assert relmax(h.numpy(), d.numpy()) < 2e-6


def test_batch64_full_size_properties():
    """Config 2 size (64 pairs of 147x147): properties that need no oracle at this size.
    (1) every pair equals the same pair run alone (no cross-talk between CTAs / atomics);
    (2) the colour maps are linear in the input images for fixed parameters;
    (3) confidence and boundary stay in [0,1], depth is 0 wherever confidence is 0."""
    S, B = 147, 64
    g = geom(S)
    est = O.restore_global(synth.raw_global(B, g.L, seed=81))
    img = planar_pair(synth.image_pairs(B, S, S, seed=82))
    ctx = _ctx(S, max_batch=B)
    full = _run_b(ctx, est, img)
    for b in (0, 17, 63):
        one = _run_b(ctx, est[b:b + 1], img[b:b + 1])
        for f, o in zip(full, one):
            assert relmax(f[b:b + 1].numpy(), o.numpy()) < 2e-6
    img2 = planar_pair(synth.image_pairs(B, S, S, seed=83))
    mix = _run_b(ctx, est, (0.25 * img + 0.75 * img2).contiguous())
    other = _run_b(ctx, est, img2)
    for k in range(3):
        assert relmax(mix[k].numpy(), (0.25 * full[k] + 0.75 * other[k]).numpy()) < 5e-6
    assert full[5].min() >= 0 and full[5].max() <= 1 + 1e-6 and full[3].min() >= 0 and full[3].max() <= 1 + 1e-6
    assert (full[4][full[5] == 0] == 0).all()
    # spot-check one pair of the big batch against the oracle
    ref = O.inference(est[5:6].to(F64), img[5:6].to(F64), g, CAM, 10.39, None)
    for name, r, o in zip(MAPS, ref, full[:6]):
        assert relmax(o[5:6].numpy(), r.numpy()) < TOL[name], name


def test_errors_are_loud():
    from blurry_edges_b200 import BlurryEdgesError, _lib
    S = GEOMS['tiny']
    ctx = _ctx(S)
    g, est, img = inference_inputs('tiny', 'normal', F32)
    with pytest.raises(BlurryEdgesError):
        ctx.render_fold(est, img.cuda(), _lib.planar_layout(S, S))           # CPU tensor
    with pytest.raises(BlurryEdgesError):
        ctx.render_fold(est.cuda().double(), img.cuda(), _lib.planar_layout(S, S))
    with pytest.raises(BlurryEdgesError):
        ctx.render_fold(torch.cat([est, est]).cuda(), torch.cat([img, img]).cuda(), _lib.planar_layout(S, S))  # B > max_batch
    assert len(ctx.render_fold(est[:0].cuda(), img[:0].cuda(), _lib.planar_layout(S, S))) == 7   # empty batch is a no-op


def test_precal_local_colors_vs_oracle_and_golden():
    """global_data_pre_cal.PostProcess equivalent: flat batch of patches, [N,R,R,3] pixels."""
    import argparse
    from blurry_edges_b200 import PostProcessLocalFused
    g = geom(147)
    est = synth.est_local(1, 40, seed=21)[0]
    pat = synth.image_pairs(40, 21, 21, seed=22)[:, 0]
    args = argparse.Namespace(R=21, w=1.0, alpha_lambda=5e-3, batch_size=1, mag=4.0,
                              cam_params={'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6})
    col = PostProcessLocalFused(args, 'cuda:0')(est.cuda(), pat.cuda()).cpu().numpy()
    assert col.shape == (40, 3, 3)
    assert relmax(col, O.colors_local(est.to(F64), pat.to(F64), g).numpy()) < 1e-5
    assert relmax(col, Golden('inference')('precal/f64')) < 1e-5


def test_repeatability_full_size_stress():
    """The warp-specialised renderer hands data between warps through named barriers, an mbarrier and cp.async; a protocol slip
    would show up as rare, run-dependent corruption.  20 launches on the same 64-pair batch must agree with the first to within the
    reordering noise of the floating-point reductions (the fold's atomics), and the discrete depth count exactly."""
    S, B = 147, 64
    g = geom(S)
    est = O.restore_global(synth.raw_global(B, g.L, seed=85)).cuda()
    img = planar_pair(synth.image_pairs(B, S, S, seed=86)).cuda()
    ctx = _ctx(S, max_batch=B)
    from blurry_edges_b200 import _lib
    lay = _lib.planar_layout(S, S)
    first = [o.clone() for o in ctx.render_fold(est, img, lay)]
    for _ in range(20):
        again = ctx.render_fold(est, img, lay)
        for name, a, b in zip(MAPS, first, again):
            err = float((a - b).abs().max() / a.abs().max())
            assert err < (1e-6 if name != 'conf' else 1e-7), (name, err)


def test_repeatability_full_size_stress_deterministic_fold_is_bit_identical():
    """The same 20 launches with the fixed-order fold (be_ctx_set_deterministic; PostProcessFused follows
    torch.use_deterministic_algorithms): every map of every launch EQUALS the first bit for bit, and the fixed-order result agrees with
    the atomic fold to reorder noise."""
    S, B = 147, 64
    g = geom(S)
    est = O.restore_global(synth.raw_global(B, g.L, seed=85, kind='stress')).cuda()
    img = planar_pair(synth.image_pairs(B, S, S, seed=86)).cuda()
    ctx = _ctx(S, max_batch=B)
    from blurry_edges_b200 import _lib
    lay = _lib.planar_layout(S, S)
    atomic = [o.clone() for o in ctx.render_fold(est, img, lay)]
    ctx.set_deterministic(True)
    first = [o.clone() for o in ctx.render_fold(est, img, lay)]
    for _ in range(20):
        again = ctx.render_fold(est, img, lay)
        for name, a, b in zip(MAPS + ('depth_thresholded',), first, again):
            assert torch.equal(a, b), name
    for name, a, b in zip(MAPS, first, atomic):
        assert float((a - b).abs().max() / a.abs().max()) < (1e-6 if name != 'conf' else 1e-7), name
    ctx.set_deterministic(False)
    back = ctx.render_fold(est, img, lay)
    assert float((back[0] - atomic[0]).abs().max()) < 1e-6


@pytest.mark.parametrize('densify', [None, 'w'])
def test_deterministic_fold_vs_oracle_small(densify):
    g, est, img = inference_inputs('mid', 'normal', F32)
    ctx = _ctx(GEOMS['mid'])
    ctx.set_deterministic(True)
    got = _run_b(ctx, est, img, densify)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, densify)
    for name, r, o in zip(MAPS, ref, got[:6]):
        assert relmax(o.numpy(), r.numpy()) < TOL[name], name
