"""GPU parity tests of the training criteria (GlobalLossFused / LocalLossFused, forward + analytic backward) against
autograd through the fp64 oracle and against the golden vectors produced by the unmodified reference.

Tolerances: loss / terms 5e-6 relative.  Gradients (north_star: 1e-5 relative in fp32): max|g - g64| / max|g64| and rel-L2 must be
<= 1e-5, OR within twice the FLOOR measured in the same test - the error of fp32 autograd through the same formulas (the oracle
run in fp32, i.e. what the reference's own fp32 path achieves; one sample of a noisy quantity, hence the factor).  Both numbers are
printed (-s) next to every assertion.  Measured (r2g): every composite loss (gamma_idx 0, final gammas, the training call, the
147x147 pair, the basic-shape scenes) is below 1e-5 outright; only single-term decompositions (one gamma at a time), stress
parameters (eta < 0.01) and the local-loss fixture (eta down to 1.6e-3) need their floor."""
import argparse

import numpy as np
import pytest
import torch

import synth
from common import F32, F64, GEOMS, Golden, geom, gloss_inputs, relmax
from oracle import be_oracle as O

pytestmark = pytest.mark.gpu
CAM = O.Camera()
CAMP = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
RANGES = dict(gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
              gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
              gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])           # utils/args.py:53-63
NAMES = ('gamma_color', 'gamma_color_cons', 'gamma_bndry_cons', 'gamma_smthns', 'gamma_smthns_cons', 'gamma_bndry_loc', 'gamma_depth')


def _gargs(S, B):
    return argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], batch_size=B, mag=4.0, cam_params=CAMP, **RANGES)


def _grad_err(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / np.abs(ref).max()), float(np.linalg.norm(got - ref) / np.linalg.norm(ref))


def _fp32_floor(raw, img_ny, img_gt, bd, deri, zgt, gam, g, g64):
    """Error of fp32 autograd through the oracle's (= the reference's) formulas against the fp64 gradient."""
    r32 = raw.clone().requires_grad_(True)
    (g32,) = torch.autograd.grad(O.global_loss(r32, img_ny, img_gt, bd, deri, zgt, gam, g, CAM), r32)
    return _grad_err(g32.numpy(), g64)


def _assert_grad(label, grad, g64, floor, factor=2.0):
    emax, el2 = _grad_err(grad, g64)
    print(f'{label}: grad err max {emax:.2e} rel-L2 {el2:.2e} | fp32-autograd floor max {floor[0]:.2e} rel-L2 {floor[1]:.2e}')
    assert emax <= max(1e-5, factor * floor[0]) and el2 <= max(1e-5, factor * floor[1]), (label, emax, el2, floor)


def _set_gammas(crit, gam):
    for n, v in zip(NAMES, gam):
        setattr(crit, n, float(v))


def _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt):
    est = raw.clone().cuda().requires_grad_(True)
    loss = crit(est, img_ny.cuda(), img_gt.cuda(), bd.cuda(), deri.cuda(), zgt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    return loss.item(), est.grad.cpu().numpy(), crit.terms.cpu().numpy()


@pytest.mark.parametrize('gname', list(GEOMS))
@pytest.mark.parametrize('gset', ['idx0', 'final'] + [f'only{k}' for k in range(7)])
def test_global_loss_vs_oracle_and_golden(gname, gset):
    from blurry_edges_b200 import GlobalLossFused
    gold = Golden('global_loss')
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs(gname, 'normal', F32)
    gam = gold(f'{gname}/gloss/normal/{gset}/f64/gammas')
    crit = GlobalLossFused(_gargs(GEOMS[gname], 2), None, 'cuda:0')
    crit.update_gamma()
    if gset == 'idx0':
        np.testing.assert_allclose(crit.gammas(), gam, rtol=0, atol=0)      # schedule mirrors global_training.py:28-51
    _set_gammas(crit, gam)
    loss, grad, terms = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    r64 = raw.to(F64).requires_grad_(True)
    l64, t64, aux = O.global_loss(r64, img_ny.to(F64), img_gt.to(F64), bd.to(F64), deri.to(F64), zgt.to(F64), gam, g, CAM, return_terms=True)
    (g64,) = torch.autograd.grad(l64, r64)
    assert abs(loss - l64.item()) <= 5e-6 * abs(l64.item())
    np.testing.assert_allclose(terms, t64.detach().numpy(), rtol=5e-6)
    floor = _fp32_floor(raw, img_ny, img_gt, bd, deri, zgt, gam, g, g64.numpy())
    # single-term decompositions (one gamma = 1, the others 0) isolate one term's gradient, whose maximum is small: 2.5 x the floor
    # (measured: smoothness alone at 45x45 is 1.6e-5 against a floor of 8.3e-6; every composite loss is below 1e-5 outright)
    factor = 2.5 if gset.startswith('only') else 2.0
    _assert_grad(f'{gname}/{gset} vs oracle', grad, g64.numpy(), floor, factor)
    assert relmax(crit.global_image.cpu().numpy(), aux['gimg'].numpy()) < 1e-5
    assert relmax(crit.global_bndry.cpu().numpy(), aux['gbnd'].numpy()) < 1e-5
    # the unmodified reference (fp64): same inputs up to the fp32 rounding of nothing (inputs are fp32-exact)
    assert abs(loss - float(gold(f'{gname}/gloss/normal/{gset}/f64/loss'))) <= 5e-6 * abs(loss)
    _assert_grad(f'{gname}/{gset} vs golden', grad, gold(f'{gname}/gloss/normal/{gset}/f64/grad'), floor, factor)


def test_global_loss_stress_parameters_within_fp32_noise_floor():
    """U[-1,1] raw parameters (eta down to 1e-3): bound = the error of fp32 autograd through the oracle itself."""
    from blurry_edges_b200 import GlobalLossFused
    gold = Golden('global_loss')
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('mid', 'stress', F32)
    gam = gold('mid/gloss/stress/idx0/f64/gammas')
    crit = GlobalLossFused(_gargs(GEOMS['mid'], 2), None, 'cuda:0')
    _set_gammas(crit, gam)
    loss, grad, _ = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    g64 = gold('mid/gloss/stress/idx0/f64/grad')
    r32 = raw.clone().requires_grad_(True)
    (g32,) = torch.autograd.grad(O.global_loss(r32, img_ny, img_gt, bd, deri, zgt, gam, g, CAM), r32)
    assert abs(loss - float(gold('mid/gloss/stress/idx0/f64/loss'))) <= 2e-5 * abs(loss)
    _assert_grad('mid/stress vs golden', grad, g64, _grad_err(g32.numpy(), g64))


@pytest.mark.parametrize('gname', list(GEOMS))
def test_global_loss_training_call_passes_the_same_tensor_twice(gname):
    """global_training.py:210 calls criteria(est, img_gt, img_gt, ...): the clean image is both the colour-solve input and the
    colour target.  The library detects the shared tensor (same device pointer) and runs the kernel variant that never reads the
    GT planes of the packed targets; the result must equal the oracle fed the same tensor twice, and the two-tensor path fed a copy."""
    from blurry_edges_b200 import GlobalLossFused
    g, raw, _, img_gt, bd, deri, zgt = gloss_inputs(gname, 'normal', F32)
    gam = [1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 0.5]
    crit = GlobalLossFused(_gargs(GEOMS[gname], 2), None, 'cuda:0')
    _set_gammas(crit, gam)
    est = raw.clone().cuda().requires_grad_(True)
    gt_d = img_gt.cuda()
    loss = crit(est, gt_d, gt_d, bd.cuda(), deri.cuda(), zgt.cuda())
    loss.backward()
    assert crit.ctx.last_same_gt is True
    terms = crit.terms.cpu().numpy()
    r64 = raw.to(F64).requires_grad_(True)
    l64, t64, _ = O.global_loss(r64, img_gt.to(F64), img_gt.to(F64), bd.to(F64), deri.to(F64), zgt.to(F64), gam, g, CAM, return_terms=True)
    (g64,) = torch.autograd.grad(l64, r64)
    assert abs(loss.item() - l64.item()) <= 5e-6 * abs(l64.item())
    np.testing.assert_allclose(terms, t64.detach().numpy(), rtol=5e-6)
    _assert_grad(f'{gname} training call', est.grad.cpu().numpy(), g64.numpy(),
                 _fp32_floor(raw, img_gt, img_gt, bd, deri, zgt, gam, g, g64.numpy()))
    # the general variant on a copy of the tensor: same numbers (the arithmetic is identical, only the loads differ)
    loss2, grad2, terms2 = _run_global(crit, raw, img_gt, img_gt.clone(), bd, deri, zgt)
    assert crit.ctx.last_same_gt is False
    assert abs(loss2 - loss.item()) <= 1e-6 * abs(loss2)
    assert relmax(grad2, est.grad.cpu().numpy()) < 1e-6


def test_global_loss_no_grad_and_eval_mode():
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('tiny', 'normal', F32)
    crit = GlobalLossFused(_gargs(GEOMS['tiny'], 2), None, 'cuda:0')
    crit.final_gamma()
    with torch.no_grad():
        l0 = crit(raw.cuda(), img_ny.cuda(), img_gt.cuda(), bd.cuda(), deri.cuda(), zgt.cuda())
    assert not l0.requires_grad
    l1, _, _ = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    assert abs(l0.item() - l1) <= 1e-7 * abs(l1)


def test_global_loss_batch_split_matches_full_batch():
    """Data-parallel semantics without NCCL: two half batches, mask counts summed between the stages, global patch
    count in the normalisers -> loss and gradients of the full batch (the depth term divides by the WHOLE batch's mask
    count, global_training.py:127)."""
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('mid', 'normal', F32, B=4)
    gam = [1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 0.5]
    crit = GlobalLossFused(_gargs(GEOMS['mid'], 4), None, 'cuda:0')
    _set_gammas(crit, gam)
    loss, grad, terms = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    dev = lambda t: t.cuda().contiguous()
    halves = [crit.__class__(_gargs(GEOMS['mid'], 2), None, 'cuda:0') for _ in range(2)]
    cnts, part = [], []
    for k, h in enumerate(halves):
        sl = slice(2 * k, 2 * k + 2)
        _, _, c = h.ctx.global_loss_stage1(dev(raw[sl]), dev(img_ny[sl]), dev(img_gt[sl]), dev(bd[sl]), dev(deri[sl]), dev(zgt[sl]))
        cnts.append(c)
    total = cnts[0] + cnts[1]
    for h in halves:
        part.append(h.ctx.global_loss_stage2(2, gam, 4 * g.L, total, True))
    # unmasked terms are means over the global batch -> partial terms add up; the masked term adds up too (same divisor)
    t_sum = (part[0][0] + part[1][0]).cpu().numpy()
    np.testing.assert_allclose(t_sum, terms, rtol=2e-6)
    assert abs((part[0][1] + part[1][1]).item() - loss) <= 2e-6 * abs(loss)
    g_cat = torch.cat([part[0][2], part[1][2]]).cpu().numpy()
    assert relmax(g_cat, grad) < 2e-6


def test_global_loss_split_stage2_equals_fused_stage2():
    """be_global_loss_stage2_launch / _finish (the loss kernel starts before the batch mask count is final, so a data-parallel
    caller can overlap the all-reduce of the count with it) against be_global_loss_stage2 on the same inputs: same terms and loss,
    gradients equal up to the rounding of grad += grad_depth / count; without gradients too; and with an empty mask (count 0)."""
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('mid', 'normal', F32, B=2)
    gam = [1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 0.5]
    crit = GlobalLossFused(_gargs(GEOMS['mid'], 2), None, 'cuda:0')
    dev = lambda t: t.cuda().contiguous()
    for z in (zgt, torch.zeros_like(zgt)):
        _, _, cnt = crit.ctx.global_loss_stage1(dev(raw), dev(img_ny), dev(img_gt), dev(bd), dev(deri), dev(z))
        t0, l0, g0 = crit.ctx.global_loss_stage2(2, gam, 2 * g.L, cnt, True)
        grad, gdep = crit.ctx.global_loss_stage2_launch(2, gam, 2 * g.L, True)
        t1, l1, g1 = crit.ctx.global_loss_stage2_finish(2, gam, 2 * g.L, cnt, grad, gdep)
        gn, gdn = crit.ctx.global_loss_stage2_launch(2, gam, 2 * g.L, False)
        assert gn is None and gdn is None
        t2, l2, _ = crit.ctx.global_loss_stage2_finish(2, gam, 2 * g.L, cnt, None, None)
        torch.cuda.synchronize()
        if int(cnt[0].item()) == 0:
            # 0/0: loss NaN and, as autograd through the reference's division, NaN for every eta coefficient (the depth term reaches
            # only those); both paths agree (the deferred path used to leave finite gradients)
            assert torch.isnan(l0).all() and torch.isnan(l1).all() and torch.isnan(l2).all()
            for gg in (g0, g1):
                assert torch.isfinite(gg[..., :8]).all() and torch.isnan(gg[..., 8:]).all()
            assert float(gdep.abs().max()) == 0.0
            np.testing.assert_array_equal(t1.cpu().numpy()[:6], t0.cpu().numpy()[:6])
            assert relmax(g1[..., :8].cpu().numpy(), g0[..., :8].cpu().numpy()) < 1e-6
        else:
            np.testing.assert_array_equal(t1.cpu().numpy(), t0.cpu().numpy())
            np.testing.assert_array_equal(t2.cpu().numpy(), t0.cpu().numpy())
            assert l1.item() == l0.item() == l2.item()
            assert float(gdep.abs().max()) > 0.0
            assert relmax(g1.cpu().numpy(), g0.cpu().numpy()) < 1e-6


def test_global_loss_full_size_one_pair_vs_oracle():
    """Config 3 geometry (147x147, 4096 patches): one pair against autograd through the fp64 oracle."""
    from blurry_edges_b200 import GlobalLossFused
    S = 147
    g = geom(S)
    img_ny = synth.image_pairs(1, S, S, seed=31)
    img_gt, bd, deri, zgt = synth.loss_targets(1, S, S, seed=31)
    raw = synth.raw_global(1, g.L, seed=33)
    gam = O.gamma_schedule(0, [RANGES[n] for n in NAMES])
    crit = GlobalLossFused(_gargs(S, 1), None, 'cuda:0')
    crit.update_gamma()
    loss, grad, terms = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    r64 = raw.to(F64).requires_grad_(True)
    l64, t64, _ = O.global_loss(r64, img_ny.to(F64), img_gt.to(F64), bd.to(F64), deri.to(F64), zgt.to(F64), gam, g, CAM, return_terms=True)
    (g64,) = torch.autograd.grad(l64, r64)
    assert abs(loss - l64.item()) <= 5e-6 * abs(l64.item())
    np.testing.assert_allclose(terms, t64.detach().numpy(), rtol=5e-6)
    _assert_grad('147x147 one pair', grad, g64.numpy(), _fp32_floor(raw, img_ny, img_gt, bd, deri, zgt, gam, g, g64.numpy()))


@pytest.mark.parametrize('name', ['final', 'loc', 'smth'])
def test_local_loss_vs_oracle_and_golden(name):
    from blurry_edges_b200 import LocalLossFused
    gl = Golden('local_loss')
    g = geom(147)
    est, ny, gt, bd, deri = synth.local_batch(8, 21, seed=41)
    betas = gl(f'lloss/{name}/f64/betas')
    args = argparse.Namespace(R=21, w=1.0, alpha_lambda=5e-3, batch_size=8, mag=4.0, cam_params=CAMP, beta_bndry_loc=0.001,
                              beta_smthns=0.0005, dynamic_epoch=200)
    crit = LocalLossFused(args, 'cuda:0')
    crit.final_beta()
    if name == 'final':
        assert (crit.beta_bndry_loc, crit.beta_smthns) == tuple(betas)
    crit.beta_bndry_loc, crit.beta_smthns = float(betas[0]), float(betas[1])
    leaf = est.clone().cuda().requires_grad_(True)
    out = leaf * 1.0                                              # network output stand-in (non-leaf, as in local_training.py:102)
    loss = crit(out, ny.cuda(), gt.cuda(), bd.cuda(), deri.cuda())
    loss.backward()
    grad = leaf.grad.cpu().numpy()
    # like the reference (local_training.py:33) the angles of the network output are wrapped in place
    assert torch.allclose(out.detach().cpu()[:, 4:8], torch.remainder(est[:, 4:8], 2 * torch.pi))
    assert abs(loss.item() - float(gl(f'lloss/{name}/f64/loss'))) <= 5e-6 * abs(loss.item())
    e64 = est.to(F64).requires_grad_(True)
    l64, t64, _ = O.local_loss(e64, ny.to(F64), gt.to(F64), bd.to(F64), deri.to(F64), betas, g, return_terms=True)
    (g64,) = torch.autograd.grad(l64, e64)
    np.testing.assert_allclose(crit.terms.cpu().numpy(), t64.detach().numpy(), rtol=5e-6)
    e32 = est.clone().requires_grad_(True)
    (g32,) = torch.autograd.grad(O.local_loss(e32, ny, gt, bd, deri, betas, g), e32)
    floor = _grad_err(g32.numpy(), g64.numpy())                   # eta reaches 1.6e-3 in this fixture: the fp32 noise floor applies
    for tag, ref in (('oracle', g64.numpy()), ('golden', gl(f'lloss/{name}/f64/grad'))):
        _assert_grad(f'local loss {name} vs {tag}', grad, ref, floor)


def test_training_repeatability_full_size_stress():
    """Same idea for the loss kernels (three CTA barriers per patch, double-buffered records, a rotating chain-rule warp): 8 repeated
    32-pair steps give the same loss (fp64 final reduction, fixed order: bit-identical up to the fold's atomics) and gradients."""
    from blurry_edges_b200 import GlobalLossFused
    S, B = 147, 32
    g = geom(S)
    crit = GlobalLossFused(_gargs(S, B), None, 'cuda:0')
    crit.update_gamma()
    raw = synth.raw_global(B, g.L, seed=300)
    img = synth.image_pairs(B, S, S, seed=301)
    gt, bd, deri, zg = synth.loss_targets(B, S, S, seed=302)
    l0, g0, _ = _run_global(crit, raw, img, gt, bd, deri, zg)
    assert np.isfinite(l0) and np.isfinite(g0).all()
    for _ in range(8):
        l1, g1, _ = _run_global(crit, raw, img, gt, bd, deri, zg)
        assert abs(l1 - l0) <= 1e-6 * abs(l0)
        assert float(np.abs(g1 - g0).max()) <= 2e-6 * float(np.abs(g0).max())


def test_global_loss_step_is_cuda_graph_capturable():
    """Both loss stages are stream-ordered with no host synchronisation (the batch mask count stays on the device), so a training
    step can be captured into a CUDA graph together with the networks around it; replays match the eager step."""
    from blurry_edges_b200 import GlobalLossFused
    S, B = GEOMS['mid'], 2
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('mid', 'normal', F32)
    crit = GlobalLossFused(_gargs(S, B), None, 'cuda:0')
    crit.update_gamma()
    dev = [t.cuda() for t in (img_ny, img_gt, bd, deri, zgt)]
    l_ref, g_ref, _ = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    est = raw.clone().cuda().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            est.grad = None
            crit(est, *dev).backward()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    est.grad = None
    with torch.cuda.graph(graph):
        loss = crit(est, *dev)
        loss.backward()
    for scale in (1.0, 0.5):
        with torch.no_grad():
            est.copy_(raw.cuda() * scale)
        graph.replay()
        torch.cuda.synchronize()
        if scale == 1.0:
            assert abs(loss.item() - l_ref) <= 1e-6 * abs(l_ref)
            assert float(np.abs(est.grad.cpu().numpy() - g_ref).max()) <= 2e-6 * float(np.abs(g_ref).max())
        else:
            l2, g2, _ = _run_global(crit, raw * scale, img_ny, img_gt, bd, deri, zgt)
            assert abs(loss.item() - l2) <= 1e-6 * abs(l2)
            assert float(np.abs(est.grad.cpu().numpy() - g2).max()) <= 2e-6 * float(np.abs(g2).max())


def test_empty_depth_mask_matches_autograd_through_the_reference_formula():
    """bndry_depth == 0 everywhere: the depth term is 0/0 (global_training.py:127).  Autograd through the reference formula leaves
    the geometric gradients finite and hands NaN to the eta coefficients of every wedge that owns a mask pixel in its patch (the
    where() of :88-90 passes exact zeros to the others).  The library - fused and deferred stage 2 alike - returns a NaN loss,
    NaN for ALL eta-coefficient gradients (a superset: the step is dead either way) and the same finite geometric gradients."""
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('tiny', 'normal', F32)
    gam = [1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 0.5]
    z0 = torch.zeros_like(zgt)
    crit = GlobalLossFused(_gargs(GEOMS['tiny'], 2), None, 'cuda:0')
    _set_gammas(crit, gam)
    loss, grad, terms = _run_global(crit, raw, img_ny, img_gt, bd, deri, z0)
    r64 = raw.to(F64).requires_grad_(True)
    l64 = O.global_loss(r64, img_ny.to(F64), img_gt.to(F64), bd.to(F64), deri.to(F64), z0.to(F64), gam, g, CAM)
    (g64,) = torch.autograd.grad(l64, r64)
    assert np.isnan(loss) and torch.isnan(l64)
    assert torch.isfinite(g64[..., :8]).all() and float(torch.isnan(g64[..., 8:]).float().mean()) > 0.5    # the reference's autograd
    assert np.isnan(grad[..., 8:]).all() and np.isfinite(grad[..., :8]).all()
    assert relmax(grad[..., :8], g64[..., :8].numpy()) < 2e-5


def test_autograd_nan_corner_case_d_equals_a_equals_zero_is_pinned():
    """utils/postprocessing_loss.py:67-78: where(a < 0, sqrt(d^2 + a^2 w^2) * sgn, d) has a NaN gradient under autograd when a pixel
    sits exactly on a wedge vertex (d = a = 0: the unselected sqrt branch contributes 0 * inf).  The centre of the fp32 pixel grid
    is linspace(-1, 1, 21)[10] = -2^-26, which the local-stage parameters (used unscaled, local_training.py:32-36) can hit exactly.
    The library does NOT reproduce the NaN: it returns the gradient with the unselected branch contributing
    exactly zero - what the reference computes once the sqrt is guarded by the same mask.  Pinned here: (1) the unguarded formula
    gives NaN for exactly the doctored patches, (2) ours is finite everywhere, (3) ours equals the guarded formula on all patches."""
    from blurry_edges_b200 import LocalLossFused
    g = geom(147)
    est, ny, gt, bd, deri = synth.local_batch(8, 21, seed=41)
    est = est.clone()
    centre = float(torch.linspace(-1, 1, 21)[10])
    assert centre == -2.0 ** -26
    doctored = [1, 5]
    for b in doctored:
        est[b, :4] = centre
    betas = (0.001, 0.0005)
    args = argparse.Namespace(R=21, w=1.0, alpha_lambda=5e-3, batch_size=8, mag=4.0, cam_params=CAMP, beta_bndry_loc=betas[0],
                              beta_smthns=betas[1], dynamic_epoch=200)
    crit = LocalLossFused(args, 'cuda:0')
    crit.final_beta()
    leaf = est.clone().cuda().requires_grad_(True)
    loss = crit(leaf * 1.0, ny.cuda(), gt.cuda(), bd.cuda(), deri.cuda())
    loss.backward()
    grad = leaf.grad.cpu().numpy()
    a64 = [t.to(F64) for t in (ny, gt, bd, deri)]
    e64 = est.to(F64).requires_grad_(True)
    (g_nan,) = torch.autograd.grad(O.local_loss(e64, *a64, betas, g), e64)
    assert torch.isnan(g_nan).any(-1).nonzero().flatten().tolist() == doctored      # (1)
    assert np.isfinite(grad).all() and np.isfinite(loss.item())                      # (2)
    plain_edge = O._edge

    def guarded_edge(px, py, ang, X, Y, w):                              # same values; sqrt only ever sees the selected branch
        sn, cs = torch.sin(ang), torch.cos(ang)
        dx, dy = X - px, Y - py
        d = -sn * dx + cs * dy
        a = cs * dx + sn * dy
        sg = torch.where(d < 0, -torch.ones_like(d), torch.ones_like(d))
        cap = torch.sqrt(torch.where(a < 0, d ** 2 + (a * w) ** 2, torch.ones_like(d))) * sg
        return torch.where(a < 0, cap, d)

    O._edge = guarded_edge
    try:
        e64 = est.to(F64).requires_grad_(True)
        l64 = O.local_loss(e64, *a64, betas, g)
        (g64,) = torch.autograd.grad(l64, e64)
        e32 = est.clone().requires_grad_(True)
        (g32,) = torch.autograd.grad(O.local_loss(e32, ny, gt, bd, deri, betas, g), e32)
    finally:
        O._edge = plain_edge
    assert torch.isfinite(g64).all()
    assert abs(loss.item() - l64.item()) <= 5e-6 * abs(l64.item())
    _assert_grad('d=a=0 corner vs guarded formula', grad, g64.numpy(), _grad_err(g32.numpy(), g64.numpy()))     # (3)


def test_uneven_shards_are_corrected_by_the_true_patch_count():
    """Data-parallel semantics without NCCL for a last batch without drop_last: shards of 1 and 2 pairs of a 3-pair batch.  Each
    shard launches its kernels with the patch count it can know (its own patches x 2 ranks); the all-reduced pair
    (mask count, true patch count) reaches only `finish`, which rescales terms, loss and gradient."""
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('mid', 'normal', F32, B=3)
    gam = [1.0, 0.2, 0.05, 0.005, 0.005, 1e-4, 0.5]
    crit = GlobalLossFused(_gargs(GEOMS['mid'], 3), None, 'cuda:0')
    _set_gammas(crit, gam)
    loss, grad, terms = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    dev = lambda t: t.cuda().contiguous()
    shards = [slice(0, 1), slice(1, 3)]
    ctxs = [crit.__class__(_gargs(GEOMS['mid'], 2), None, 'cuda:0').ctx for _ in shards]
    cnts = [c.global_loss_stage1(dev(raw[sl]), dev(img_ny[sl]), dev(img_gt[sl]), dev(bd[sl]), dev(deri[sl]), dev(zgt[sl]))[2]
            for c, sl in zip(ctxs, shards)]
    total = cnts[0] + cnts[1]                                            # the 16-byte all-reduce
    assert int(total[1].item()) == 3 * g.L
    parts = []
    for c, sl in zip(ctxs, shards):
        nb = sl.stop - sl.start
        assumed = nb * g.L * 2                                           # local patches x world: what sync_loss_normalisers returns
        gr, gd = c.global_loss_stage2_launch(nb, gam, assumed, True)
        parts.append(c.global_loss_stage2_finish(nb, gam, assumed, total, gr, gd))
    np.testing.assert_allclose((parts[0][0] + parts[1][0]).cpu().numpy(), terms, rtol=2e-6)
    assert abs((parts[0][1] + parts[1][1]).item() - loss) <= 2e-6 * abs(loss)
    assert relmax(torch.cat([parts[0][2], parts[1][2]]).cpu().numpy(), grad) < 2e-6


def test_host_buffer_training_entry_matches_the_device_path():
    """be_host_global_loss (host buffers in, terms / loss / grad out; chunked H2D overlapped with the kernels, depth normaliser
    deferred) against GlobalLossFused on device tensors, for the training call (one image passed twice) and the validation call."""
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('mid', 'normal', F32, B=5)
    crit = GlobalLossFused(_gargs(GEOMS['mid'], 5), None, 'cuda:0')
    crit.update_gamma()
    for first in (img_gt, img_ny):
        loss, grad, terms = _run_global(crit, raw, first, img_gt, bd, deri, zgt)
        ht, hl, hg = crit.ctx.host_global_loss(raw, first, img_gt, bd, deri, zgt, crit.gammas())
        np.testing.assert_allclose(ht.numpy(), terms, rtol=2e-6)
        assert abs(hl.item() - loss) <= 2e-6 * abs(loss)
        assert relmax(hg.numpy(), grad) < 2e-6
        ht2, hl2, none = crit.ctx.host_global_loss(raw, first, img_gt, bd, deri, zgt, crit.gammas(), want_grad=False)
        assert none is None and hl2.item() == hl.item()


@pytest.mark.parametrize('B', [1, 2, 3, 7])
def test_host_buffer_training_entry_both_schedules_any_batch(B):
    """The two schedules of the host-buffer step - two-phase (single process: final count before the loss kernels, gradient home per
    chunk) and deferred (begin / end, what data-parallel callers use) - on batches smaller and larger than the number of chunks: both
    equal the device-resident path."""
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('tiny', 'normal', F32, B=B)
    crit = GlobalLossFused(_gargs(GEOMS['tiny'], B), None, 'cuda:0')
    crit.update_gamma()
    loss, grad, terms = _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    for deferred in (False, True):
        for _ in range(2):                                        # twice: the second call reuses staging buffers, events and streams
            ht, hl, hg = crit.ctx.host_global_loss(raw, img_ny, img_gt, bd, deri, zgt, crit.gammas(), deferred=deferred)
            np.testing.assert_allclose(ht.numpy(), terms, rtol=2e-6)
            assert abs(hl.item() - loss) <= 2e-6 * abs(loss)
            assert relmax(hg.numpy(), grad) < 2e-6, deferred


def test_train_timing_hook_reports_the_seven_device_operations():
    from blurry_edges_b200 import GlobalLossFused
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('mid', 'normal', F32)
    crit = GlobalLossFused(_gargs(GEOMS['mid'], 2), None, 'cuda:0')
    crit.update_gamma()
    crit.ctx.set_timing(True)
    _run_global(crit, raw, img_ny, img_gt, bd, deri, zgt)
    ms = crit.ctx.last_train_timing()
    assert len(ms) == 7 and all(0.0 <= m < 50.0 for m in ms) and all(ms[i] > 0.0 for i in (1, 2, 4, 5)), ms
    crit.ctx.set_timing(False)


@pytest.mark.parametrize('call', ['train', 'val'])
def test_basic_shape_scenes_full_size_vs_reference_golden(call):
    """Config 3 inputs as SURVEY 8(d) prescribes them: two 147x147 scenes of the reference's own generator (hard edges, flat regions, a
    true distance field, Sobel-response derivative maps, boundary depth; tests/golden/make_shapes.py) - loss and gradient of the
    unmodified GlobalLoss (fp64) for the training call (clean image twice) and the validation call."""
    from blurry_edges_b200 import GlobalLossFused
    z = synth.shapes_arrays()['z']
    S, B = 147, 2
    g = geom(S)
    ny, gt, bd, deri, zg = synth.shapes_batch(B)
    raw = synth.raw_global(B, g.L, seed=81)
    crit = GlobalLossFused(_gargs(S, B), None, 'cuda:0')
    crit.update_gamma()
    np.testing.assert_allclose(crit.gammas(), z['gammas'], rtol=0, atol=0)
    first = gt if call == 'train' else ny
    est = raw.clone().cuda().requires_grad_(True)
    f_d, gt_d = first.cuda(), gt.cuda()
    loss = crit(est, gt_d if call == 'train' else f_d, gt_d, bd.cuda(), deri.cuda(), zg.cuda())
    loss.backward()
    g_ref = z[f'{call}.grad']
    assert abs(loss.item() - float(z[f'{call}.loss'])) <= 5e-6 * abs(loss.item())
    emax, el2 = _grad_err(est.grad.cpu().numpy(), g_ref)
    fmax, fl2 = _grad_err(z['train32.grad'], z['train.grad'])             # the unmodified reference's own fp32 run against its fp64 run
    print(f'shapes {call}: grad err max {emax:.2e} rel-L2 {el2:.2e}; reference fp32 floor max {fmax:.2e} rel-L2 {fl2:.2e}')
    assert emax < 1e-5 and el2 < 1e-5, (emax, el2, fmax, fl2)          # realistic inputs: the north_star bound outright


def test_deterministic_fold_is_bit_identical_over_20_launches():
    """global_training.py:177 -> set_seed(1898, deterministic=True) -> torch.use_deterministic_algorithms(True)
    (utils/util_func.py:17-19).  GlobalLossFused follows that switch: the fold then runs in fixed order (one CTA per patch row writing a
    private slab, slabs added in ascending patch-row order) and 20 full-size steps on stress parameters are BIT-identical in loss,
    gradient and folded maps; the atomic fold agrees with it to reorder noise."""
    from blurry_edges_b200 import GlobalLossFused
    S, B = 147, 8
    g = geom(S)
    crit = GlobalLossFused(_gargs(S, B), None, 'cuda:0')
    crit.update_gamma()
    raw = synth.raw_global(B, g.L, seed=300, kind='stress')
    img = synth.image_pairs(B, S, S, seed=301)
    gt, bd, deri, zg = synth.loss_targets(B, S, S, seed=302)
    l_at, g_at, _ = _run_global(crit, raw, img, gt, bd, deri, zg)
    gi_at = crit.global_image.clone()
    torch.use_deterministic_algorithms(True)
    try:
        l0, g0, _ = _run_global(crit, raw, img, gt, bd, deri, zg)
        gi0, gb0 = crit.global_image.clone(), crit.global_bndry.clone()
        for _ in range(20):
            l1, g1, _ = _run_global(crit, raw, img, gt, bd, deri, zg)
            assert l1 == l0 and np.array_equal(g1, g0)
            assert torch.equal(crit.global_image, gi0) and torch.equal(crit.global_bndry, gb0)
    finally:
        torch.use_deterministic_algorithms(False)
    assert abs(l_at - l0) <= 1e-6 * abs(l0) and float(np.abs(g_at - g0).max()) <= 2e-6 * float(np.abs(g0).max())
    assert float((gi_at - gi0).abs().max()) <= 2e-6
    l2, _, _ = _run_global(crit, raw, img, gt, bd, deri, zg)                 # the switch is followed both ways
    assert abs(l2 - l0) <= 1e-6 * abs(l0)
