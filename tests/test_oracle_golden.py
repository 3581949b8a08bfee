"""Pin oracle/be_oracle.py to outputs of the unmodified reference (tests/golden/*.npz, made by
tests/golden/make_golden.py).  CPU only.  fp64 oracle vs fp64 reference must agree to rounding;
the reference's own fp32 result is held to its documented noise floor."""
import numpy as np
import pytest
import torch

import synth
from common import F32, F64, GEOMS, MAPS, Golden, geom, gloss_inputs, inference_inputs, planar_pair, relmax
from oracle import be_oracle as O

CAM = O.Camera()


@pytest.fixture(scope='module')
def ginf():
    return Golden('inference')


@pytest.fixture(scope='module')
def gglo():
    return Golden('global_loss')


@pytest.mark.parametrize('gname', list(GEOMS))
def test_pass_a_colors(ginf, gname):
    S = GEOMS[gname]
    g = geom(S)
    img = planar_pair(synth.image_pairs(1, S, S, seed=3, dtype=F64))[0]      # [2,3,H,W]
    est = synth.est_local(2, g.L, seed=5, dtype=F64)
    ours = O.colors_only(est, img, g).numpy()
    assert relmax(ours, ginf(f'{gname}/passA/f64')) < 1e-10
    # the reference's fp32 path (trace-formula inverse) against its own fp64: noise floor, recorded not required
    assert relmax(ginf(f'{gname}/passA/f32'), ginf(f'{gname}/passA/f64')) < 2e-2


@pytest.mark.parametrize('gname', list(GEOMS))
@pytest.mark.parametrize('densify', [None, 'w'])
@pytest.mark.parametrize('kind', ['normal', 'stress'])
def test_pass_b_maps(ginf, gname, densify, kind):
    g, est, img = inference_inputs(gname, kind, F64)
    ours = O.inference(est, img, g, CAM, 10.39, densify)
    for name, o in zip(MAPS, ours):
        ref = ginf(f'{gname}/passB/{densify or "none"}/{kind}/f64/{name}')
        assert o.shape == ref.shape, name
        assert relmax(o.numpy(), ref) < 1e-9, name


def test_precal_colors(ginf):
    g = geom(147)
    est = synth.est_local(1, 40, seed=21, dtype=F64)[0]
    pat = synth.image_pairs(40, 21, 21, seed=22, dtype=F64)[:, 0]
    assert relmax(O.colors_local(est, pat, g).numpy(), ginf('precal/f64')) < 1e-10


@pytest.mark.parametrize('gname', list(GEOMS))
@pytest.mark.parametrize('kind,gset', [('normal', 'idx0'), ('normal', 'final'), ('stress', 'idx0')]
                         + [('normal', f'only{k}') for k in range(7)])
def test_global_loss_and_grad(gglo, gname, kind, gset):
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs(gname, kind, F64)
    key = f'{gname}/gloss/{kind}/{gset}/f64'
    gammas = gglo(f'{key}/gammas')
    raw = raw.requires_grad_(True)
    loss, terms, aux = O.global_loss(raw, img_ny, img_gt, bd, deri, zgt, gammas, g, CAM, return_terms=True)
    (grad,) = torch.autograd.grad(loss, raw)
    assert abs(loss.item() - float(gglo(f'{key}/loss'))) <= 1e-11 * max(1.0, abs(loss.item()))
    assert relmax(grad.numpy(), gglo(f'{key}/grad')) < 1e-8
    if gset == 'idx0':
        assert relmax(aux['gimg'].numpy(), gglo(f'{key}/global_image')) < 1e-10
        assert relmax(aux['gbnd'].numpy(), gglo(f'{key}/global_bndry')) < 1e-10


def test_gamma_schedule_matches_reference_idx0(gglo):
    ranges = [[1.0, 0.1, 0.1], [0.2, 0.1, 0.05], [0.05, 0.05, 0.02], [0.005, 0.1, 0.002], [0.005, 0.1, 0.002],
              [0.0001, 0.05, 0.0001], [0.0001, 0.05, 0.5]]                   # utils/args.py:53-59
    np.testing.assert_allclose(O.gamma_schedule(0, ranges), gglo('tiny/gloss/normal/idx0/f64/gammas'), rtol=0, atol=0)
    np.testing.assert_allclose([r[-1] for r in ranges], gglo('tiny/gloss/normal/final/f64/gammas'), rtol=0, atol=0)


@pytest.mark.parametrize('name', ['final', 'loc', 'smth'])
def test_local_loss_and_grad(name):
    gl = Golden('local_loss')
    g = geom(147)
    est, ny, gt, bd, deri = synth.local_batch(8, 21, seed=41, dtype=F64)
    betas = gl(f'lloss/{name}/f64/betas')
    leaf = est.clone().requires_grad_(True)
    loss = O.local_loss(leaf, ny, gt, bd, deri, betas, g)
    (grad,) = torch.autograd.grad(loss, leaf)
    assert abs(loss.item() - float(gl(f'lloss/{name}/f64/loss'))) <= 1e-11
    assert relmax(grad.numpy(), gl(f'lloss/{name}/f64/grad')) < 1e-8


def test_cover_count_closed_form():
    for S in (29, 45, 147):
        g = geom(S)
        ones = torch.ones(g.L, 1, g.R, g.R, dtype=F64)
        assert torch.equal(O.fold_sum(ones, 1, g)[0, 0], O.cover_count(g, F64))
    assert O.cover_count(geom(147)).max().item() == 121 and O.cover_count(geom(147)).min().item() == 1


def test_reference_fp32_noise_floor_is_what_survey_says(ginf, gglo):
    """Documents (does not gate on) how far the reference's own fp32 path is from its fp64 path."""
    worst = 0.0
    for name in ('image', 'sharp', 'refoc'):
        k = f'mid/passB/none/normal'
        worst = max(worst, relmax(ginf(f'{k}/f32/{name}'), ginf(f'{k}/f64/{name}')))
    assert 1e-6 < worst < 5e-2   # ~1e-3: the trace-formula inverse, SURVEY section 7 hard part 1


def _eval_depth(pred, gt, msk, crop=10, tau=1.25, z_min=0.75, z_max=1.18):
    """Restatement of utils/metrics.py:3-21 (delta1..3, RMSE cm, AbsRel cm) for the metric-level check."""
    pred = np.clip(pred, z_min, z_max)[:, crop:-crop, crop:-crop]
    gt = gt[:, crop:-crop, crop:-crop]
    msk = msk[:, crop:-crop, crop:-crop]
    err = np.abs(gt - pred)
    pn = np.clip((pred - z_min) / (z_max - z_min), 0, 1)
    gn = np.clip((gt - z_min) / (z_max - z_min), 0, 1)
    n = msk.sum()
    acc = np.maximum(gn / (pn + 1e-8), pn / (gn + 1e-8))
    d = [np.sum((acc < tau ** k) * msk) / n for k in (1, 2, 3)]
    return (*d, np.sqrt(np.sum(err ** 2 * msk) / n) * 100, np.sum(err * msk / gt * msk) / n * 100)


@pytest.mark.parametrize('densify', [None, 'w'])
def test_config1_147_against_unchanged_driver(densify):
    """blurry_edges_test.depth_estimator (unchanged, fp32, random-init nets) vs the fp64 oracle fed the
    same `est`: maps within the reference's fp32 noise, depth metrics equal to 4 decimals."""
    gc = Golden('config1')
    S = 147
    g = geom(S)
    img32 = torch.from_numpy(synth.photon_pairs(1, S, S, seed=51, alpha=190)).float() / 190.0
    est = torch.from_numpy(gc('config1/est')).to(F64)
    maps = O.inference(est, planar_pair(img32.to(F64)), g, CAM, 10.39, densify)
    key = f'config1/{densify or "none"}'
    st = 6
    image = maps[0][0].permute(0, 2, 3, 1).numpy()[:, ::st, ::st]
    assert relmax(image, gc(f'{key}/image')) < 5e-3
    assert relmax(maps[1][0].permute(1, 2, 0).numpy()[::st, ::st], gc(f'{key}/sharp')) < 5e-3
    assert relmax(maps[2][0].permute(1, 2, 0).numpy()[::st, ::st], gc(f'{key}/refoc')) < 5e-3
    assert relmax(maps[3][0, 0].numpy()[::st, ::st], gc(f'{key}/bndry')) < 1e-4
    assert relmax(maps[5][0].numpy()[::st, ::st], gc(f'{key}/conf')) < 1e-6
    thres = 0.0 if densify == 'w' else 0.05
    depth = np.where(maps[5].numpy() > thres, maps[4].numpy(), 0.0)
    assert relmax(depth[0, ::st, ::st], gc(f'{key}/depth_thresholded')) < 1e-5
    gt = synth.uniform((1, S, S), 52, 0.75, 1.18).numpy()
    ours = _eval_depth(depth, gt.astype(np.float64), depth > 0)
    np.testing.assert_allclose(ours, gc(f'{key}/metrics'), rtol=0, atol=5e-5)


def test_big_235_stitch_against_unchanged_driver():
    """blurry_edges_test_big.depth_estimator (unchanged, fp32, stub nets) vs oracle block stitch."""
    gb = Golden('big')
    S = 235
    g = geom(147)
    img = torch.from_numpy(synth.photon_pairs(1, S, S, seed=61)).float() / 190.0
    nblk = int(gb('big235/nblocks'))
    assert nblk == 4
    est = torch.stack([O.restore_global(synth.raw_global(1, g.L, seed=70 + k, kind='normal'))[0] for k in range(nblk)])
    out = O.inference_big(est.to(F64), planar_pair(img.to(F64))[0], g, CAM, S, S)
    st = 3
    assert relmax(out[0][0].permute(0, 2, 3, 1).numpy()[:, ::st, ::st], gb('big235/image')) < 5e-3
    assert relmax(out[1][0].permute(1, 2, 0).numpy()[::st, ::st], gb('big235/sharp')) < 5e-3
    assert relmax(out[2][0].permute(1, 2, 0).numpy()[::st, ::st], gb('big235/refoc')) < 5e-3
    assert relmax(out[3][0, 0].numpy()[::st, ::st], gb('big235/bndry')) < 1e-4
    assert relmax(out[5][0].numpy()[::st, ::st], gb('big235/conf')) < 1e-6
    assert relmax(out[6][0].numpy()[::st, ::st], gb('big235/depth_thresholded')) < 1e-5


def test_oracle_on_basic_shape_scenes_of_the_reference_generator():
    """tests/golden/shapes147.npz: two full-size scenes of train_val_data_generator.SyntheticShapeDataGenerator with the golden
    loss / gradient of the unmodified GlobalLoss (training call: the clean image passed twice, global_training.py:210)."""
    import synth
    z = synth.shapes_arrays()['z']
    g = O.Geometry(H=147, W=147)
    ny, gt, bd, deri, zg = synth.shapes_batch(2, dtype=torch.float64)
    assert float(bd.max()) > 10 and float((zg != 0).float().mean()) > 0.02 and float((deri == 0).float().mean()) > 0.5   # a real scene
    raw = synth.raw_global(2, g.L, seed=81, dtype=torch.float64).requires_grad_(True)
    loss = O.global_loss(raw, gt, gt, bd, deri, zg, z['gammas'], g, O.Camera())
    (grad,) = torch.autograd.grad(loss, raw)
    assert abs(loss.item() - float(z['train.loss'])) <= 1e-10 * abs(loss.item())
    assert float((grad - torch.from_numpy(z['train.grad'])).abs().max()) <= 1e-9 * float(np.abs(z['train.grad']).max())


def test_oracle_inference_on_a_generator_scene():
    """tests/golden/shapes147_infer.npz: pass A and pass B of the unmodified blurry_edges_test.PostProcess (fp64) on the noisy pair of a
    full-size basic-shape scene - hard edges, flat regions, the generator's noise model - under both mask rules."""
    from common import shapes_inference_inputs
    gold, g, est, est10, img = shapes_inference_inputs(F64)
    assert relmax(O.colors_only(est10, img[0], g).numpy(), gold('passA')) < 1e-10
    for densify in (None, 'w'):
        ours = O.inference(est, img, g, CAM, 10.39, densify)
        for name, o in zip(MAPS, ours):
            key = f'{densify or "none"}/{name}' if name in ('refoc', 'depth', 'conf') else name
            ref = gold(key)
            assert o.shape == ref.shape, name
            assert relmax(o.numpy(), ref) < 1e-9, (densify, name)
