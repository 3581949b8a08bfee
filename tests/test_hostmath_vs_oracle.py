"""The fp32 arithmetic the CUDA kernels use (csrc/be_math.cuh) compiled for the host and checked against the fp64
oracle - the closest thing to a kernel parity test that runs without a GPU.  Tolerances are the ones the GPU parity
tests use (see tests/test_gpu_inference.py for the derivation)."""
import numpy as np
import pytest
import torch

import synth
from common import F32, F64, GEOMS, MAPS, geom, inference_inputs, planar_pair, relmax
from oracle import be_oracle as O
from oracle import hostmath

CAM = O.Camera()


@pytest.mark.parametrize('gname', list(GEOMS))
def test_pass_a(gname):
    S = GEOMS[gname]
    g = geom(S)
    img = planar_pair(synth.image_pairs(1, S, S, seed=3))[0]
    est = synth.est_local(2, g.L, seed=5)
    ref = O.colors_only(est.to(F64), img.to(F64), g).numpy()
    got = hostmath.colors(est, img, g, CAM)
    assert relmax(got, ref) < 2e-5


@pytest.mark.parametrize('gname', list(GEOMS))
@pytest.mark.parametrize('densify', [None, 'w'])
@pytest.mark.parametrize('kind,tol', [('normal', 1e-5), ('stress', 5e-4)])
def test_pass_b(gname, densify, kind, tol):
    g, est, img = inference_inputs(gname, kind, F32)
    ref = O.inference(est.to(F64), img.to(F64), g, CAM, 10.39, densify)
    got = hostmath.render_fold(est, img, g, CAM, 10.39, densify)
    for name, r, o in zip(MAPS, ref, got):
        e = relmax(o, r.numpy())
        lim = tol if name not in ('sharp',) else max(tol, 2e-3)   # eta=1e-4 render: edge pixels amplify 1-ulp distance noise
        assert e < lim, (name, e)


def _grad_err(got, ref):
    """max-norm and rel-L2 error of a gradient tensor"""
    ref = np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / np.abs(ref).max()), float(np.linalg.norm(got - ref) / np.linalg.norm(ref))


GAMMA_SETS = [[1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 1e-4], [0.1, 0.05, 0.02, 0.002, 0.002, 1e-4, 0.5]] + \
             [[1.0 if k == j else 0.0 for k in range(7)] for j in range(7)]


@pytest.mark.parametrize('gname', list(GEOMS))
@pytest.mark.parametrize('gammas', GAMMA_SETS)
def test_global_loss_forward_backward(gname, gammas):
    """Analytic backward of the kernels (be_math.cuh) vs autograd through the fp64 oracle, term by term."""
    from common import gloss_inputs
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs(gname, 'normal', F32)
    r64 = raw.to(F64).requires_grad_(True)
    loss, terms, aux = O.global_loss(r64, img_ny.to(F64), img_gt.to(F64), bd.to(F64), deri.to(F64), zgt.to(F64), gammas, g, CAM,
                                     return_terms=True)
    (gref,) = torch.autograd.grad(loss, r64)
    l, t, grad, gimg, gbnd = hostmath.global_loss(raw, img_ny, img_gt, bd, deri, zgt, gammas, g, CAM)
    assert abs(l - loss.item()) <= 2e-6 * abs(loss.item())
    np.testing.assert_allclose(t, terms.detach().numpy(), rtol=5e-6)
    assert relmax(gimg, aux['gimg'].numpy()) < 1e-5 and relmax(gbnd, aux['gbnd'].numpy()[:, 0]) < 1e-5
    emax, el2 = _grad_err(grad, gref.numpy())
    assert emax < 5e-5 and el2 < 2e-5, (emax, el2)   # fp32 sums of <=882 terms; the reference's own fp32 grads: rel-L2 1e-4


@pytest.mark.parametrize('betas', [(0.001, 0.0005), (1.0, 0.0), (0.0, 1.0)])
@pytest.mark.parametrize('eta_lo', [0.0, -0.5])
def test_local_loss_forward_backward(betas, eta_lo):
    """eta_lo = 0: eta >= 0.016 (realistic), tight bound.  eta_lo = -0.5: eta down to 1.6e-3, where a 1-ulp distance rounding
    moves the Sobel of a near-step edge by 1e-4 relative: there the bound is the fp32 noise floor itself, measured as the
    error of the oracle's own fp32 autograd against its fp64 autograd."""
    g = geom(147)
    est, ny, gt, bd, deri = synth.local_batch(8, 21, seed=41)
    est[:, 8:] = synth.uniform((8, 2), 43, eta_lo, 1.0)
    e64 = est.to(F64).requires_grad_(True)
    loss, terms, _ = O.local_loss(e64, ny.to(F64), gt.to(F64), bd.to(F64), deri.to(F64), betas, g, return_terms=True)
    (gref,) = torch.autograd.grad(loss, e64)
    l, t, grad = hostmath.local_loss(est, ny, gt, bd, deri, betas, g, CAM)
    assert abs(l - loss.item()) <= 2e-6 * abs(loss.item())
    np.testing.assert_allclose(t, terms.detach().numpy(), rtol=5e-6)
    emax, el2 = _grad_err(grad, gref.numpy())
    if eta_lo >= 0:
        assert emax < 2e-5 and el2 < 2e-5, (emax, el2)
    else:
        e32 = est.clone().requires_grad_(True)
        (g32,) = torch.autograd.grad(O.local_loss(e32, ny, gt, bd, deri, betas, g), e32)
        floor, _ = _grad_err(g32.numpy(), gref.numpy())
        assert emax < max(2e-5, 2 * floor), (emax, floor)


@pytest.mark.parametrize('seed,scale', [(0, 0.1), (1, 1.0), (2, 1.0)])
def test_packed_two_pixel_functions_equal_the_scalar_specification(seed, scale):
    """be_pack.cuh (FFMA2 halves on the GPU, plain fp32 pairs on the host) against be_math.cuh: the forward functions are the
    same IEEE operations in the same order (bit-identical), the wedge backward is algebraically rewritten (1e-5 relative)."""
    g = O.Geometry(H=21, W=21)
    raw = synth.raw_global(1, 1, seed=40 + seed) * (scale / 0.1)
    p = O.restore_global(raw)[0, 0].numpy()
    rng = np.random.default_rng(seed)
    xy = np.concatenate([np.stack(np.meshgrid(np.linspace(-1, 1, 21), np.linspace(-1, 1, 21)), -1).reshape(-1, 2),
                         rng.uniform(-1, 1, size=(559, 2))]).astype(np.float32)
    err = hostmath.pack_selfcheck(p, O.Camera(), xy)
    assert err[0] == 0.0 and err[1] == 0.0 and err[2] == 0.0 and err[3] == 0.0, err
    assert err[4] < 1e-5, err
