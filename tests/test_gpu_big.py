"""GPU tests of the blocked big-image path (config 4) against the oracle's restatement of blurry_edges_test_big.py and the
golden maps captured from the unmodified big-image driver."""
import argparse

import numpy as np
import pytest
import torch

import synth
from common import F32, F64, Golden, geom, planar_pair, relmax
from oracle import be_oracle as O

pytestmark = pytest.mark.gpu
CAM = O.Camera()
CAMP = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}
TOL = [1e-5, 2e-5, 1e-5, 1e-5, 1e-5, 1e-6]


def _args(big):
    return argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[147, 147], big_img_size=[big, big], batch_size=1,
                              mag=4.0, rho_prime=10.39, densify=None, n_margin_patch=10, cam_params=CAMP)


def _inputs(big, seed):
    g = geom(147)
    img = planar_pair(torch.from_numpy(synth.photon_pairs(1, big, big, seed=seed)).float() / 190.0)[0]      # [2,3,big,big]
    return g, img


def test_big_235_vs_golden_reference_driver():
    """Same stub-network parameters the unmodified blurry_edges_test_big.depth_estimator consumed."""
    from blurry_edges_b200 import BigImageFused
    gb = Golden('big')
    big = 235
    g, img = _inputs(big, 61)
    est = torch.stack([O.restore_global(synth.raw_global(1, g.L, seed=70 + k))[0] for k in range(4)])
    helper = BigImageFused(_args(big), None, 'cuda:0')
    assert helper.nblk == int(gb('big235/nblocks'))
    out = [o.cpu() for o in helper(est.cuda(), img.cuda())]
    st = 3
    # the golden maps are the reference's fp32 results (trace-formula inverse): its own noise floor applies
    assert relmax(out[0][0].permute(0, 2, 3, 1).numpy()[:, ::st, ::st], gb('big235/image')) < 5e-3
    assert relmax(out[3][0, 0].numpy()[::st, ::st], gb('big235/bndry')) < 1e-4
    assert relmax(out[5][0].numpy()[::st, ::st], gb('big235/conf')) < 1e-6
    assert relmax(out[6][0].numpy()[::st, ::st], gb('big235/depth_thresholded')) < 1e-5
    ref = O.inference_big(est.to(F64), img.to(F64), g, CAM, big, big)
    for k in range(6):
        assert relmax(out[k].numpy(), ref[k].numpy()) < TOL[k], k


def test_big_323_all_window_cases_and_block_sharding():
    """3x3 blocks: corner, edge and fully interior windows; two partial accumulators (a 2-rank sharding) sum to the result."""
    from blurry_edges_b200 import BigImageFused, shard_blocks
    big = 323
    g, img = _inputs(big, 63)
    est = torch.stack([O.restore_global(synth.raw_global(1, g.L, seed=170 + k))[0] for k in range(9)])
    helper = BigImageFused(_args(big), None, 'cuda:0')
    assert helper.nblk == 9
    out = [o.cpu() for o in helper(est.cuda(), img.cuda())]
    ref = O.inference_big(est.to(F64), img.to(F64), g, CAM, big, big)
    for k in range(6):
        assert relmax(out[k].numpy(), ref[k].numpy()) < TOL[k], k
    accs = []
    for r in range(2):
        lo, hi = shard_blocks(9, r, 2)
        accs.append(helper.render_partial(est[lo:hi].cuda(), img.cuda(), lo, hi))
    out2 = [o.cpu() for o in helper.finish(accs[0] + accs[1])]
    for a, b in zip(out, out2):
        assert relmax(b.numpy(), a.numpy()) < 2e-6


def test_big_pass_a_colors_per_block():
    from blurry_edges_b200 import BigImageFused
    big = 235
    g, img = _inputs(big, 61)
    helper = BigImageFused(_args(big), None, 'cuda:0')
    params = torch.stack([synth.est_local(2, g.L, seed=60 + k) for k in range(4)])                          # [4,2,L,10]
    col = helper.colors(params.cuda(), img.cuda()).cpu()
    for k, w in enumerate(helper.windows):
        blk = img[:, :, w[2]:w[2] + 147, w[3]:w[3] + 147]
        ref = O.colors_only(params[k].to(F64), blk.to(F64), g).numpy()
        assert relmax(col[k].numpy(), ref) < 1e-5


def test_big_1027_full_size_properties():
    """Config 4 size (121 blocks, 504x504 patches): sharded == unsharded, bounded maps, every pixel covered."""
    from blurry_edges_b200 import BigImageFused, shard_blocks
    big = 1027
    g, img = _inputs(big, 65)
    helper = BigImageFused(_args(big), None, 'cuda:0')
    assert helper.nblk == 121
    est = O.restore_global(synth.raw_global(121, g.L, seed=270)).cuda()
    out = helper(est, img.cuda())
    acc = None
    for r in range(8):
        lo, hi = shard_blocks(121, r, 8)
        part = helper.render_partial(est[lo:hi], img.cuda(), lo, hi)
        acc = part if acc is None else acc + part
    out8 = helper.finish(acc)
    for a, b in zip(out, out8):
        assert relmax(b.cpu().numpy(), a.cpu().numpy()) < 2e-6
    conf, bnd = out[5].cpu(), out[3].cpu()
    assert conf.min() >= 0 and conf.max() <= 1 + 1e-6 and bnd.min() >= 0 and bnd.max() <= 1 + 1e-6
    assert torch.isfinite(out[0]).all() and (out[4][out[5] == 0] == 0).all()
    # one interior block region against the oracle: the 44x44 interior patches of block (5,5) fully determine the pixels
    # whose covering patches all lie inside that window
    iv = ih = 5
    w = helper.windows[iv * 11 + ih]
    r = O.inference(est[iv * 11 + ih:iv * 11 + ih + 1].cpu().to(F64), img[None, :, :, w[2]:w[2] + 147, w[3]:w[3] + 147].to(F64), g, CAM,
                    10.39, None, return_patches=True)
    # pixel rows/cols covered only by local patches 10..53: y in [10*2+20, 54*2) -> [40, 108)
    P1 = r['P1'].reshape(64, 64, 3, 21, 21)[10:54, 10:54]
    sub = O.Geometry(H=44 * 2 + 19, W=44 * 2 + 19)
    full = O.fold_sum(P1.reshape(-1, 3, 21, 21), 1, sub)[0] / O.cover_count(sub, F64)
    ours = out[0][0, 0, :, w[2] + 20:w[2] + 20 + sub.H, w[3] + 20:w[3] + 20 + sub.W].cpu().double()
    inner = slice(20, sub.H - 20)
    assert relmax(ours[:, inner, inner].numpy(), full[:, inner, inner].numpy()) < 1e-5
