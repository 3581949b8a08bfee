"""The inference driver body (blurry_edges_test.py:114-149) on device: glue kernels + passes A/B against a restatement
of the driver built from torch ops and the oracle, with small stand-in networks (the real CNN/transformer are out of scope)."""
import argparse

import numpy as np
import pytest
import torch

import synth
from common import F32, F64, GEOMS, geom, planar_pair, relmax
from oracle import be_oracle as O
from test_oracle_golden import _eval_depth

pytestmark = pytest.mark.gpu
CAM = O.Camera()
CAMP = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}


class TinyLocal(torch.nn.Module):          # [N,3,21,21] -> [N,10]
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(3 * 9, 10)
        with torch.no_grad():
            self.lin.weight.copy_(synth.uniform((10, 27), 91, -1.5, 1.5))
            self.lin.bias.copy_(torch.tensor([0.1, -0.2, 0.3, 0.2, 1.0, 2.0, 4.0, 1.5, 0.3, 0.2]))

    def forward(self, v):
        return self.lin(torch.nn.functional.adaptive_avg_pool2d(v, 3).flatten(1))


class TinyGlobal(torch.nn.Module):         # [B,L,38] -> [B,L,12]
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(38, 12)
        with torch.no_grad():
            self.lin.weight.copy_(synth.uniform((12, 38), 92, -0.08, 0.08))
            self.lin.bias.zero_()

    def forward(self, pm):
        return self.lin(pm)


@pytest.mark.parametrize('densify', [None, 'w'])
def test_driver_body_matches_restated_reference_driver(densify):
    from blurry_edges_b200 import DepthEstimatorFused
    S, B = GEOMS['mid'], 2
    g = geom(S)
    args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], batch_size=B, mag=4.0, rho_prime=10.39,
                              densify=densify, crop=10, cam_params=CAMP)
    local_m, global_m = TinyLocal(), TinyGlobal()
    img = synth.image_pairs(B, S, S, seed=95)                                   # [B,2,H,W,3]
    gt = synth.uniform((B, S, S), 96, 0.75, 1.18)
    est = DepthEstimatorFused(args, local_m.cuda(), global_m.cuda(), 'cuda:0')
    res = est(img.cuda(), gt.cuda())
    # ---- restated driver (blurry_edges_test.py:119-149), fp64 oracle for the path, same stand-in networks in fp32 ----
    local_m, global_m = local_m.cpu(), global_m.cpu()
    maps, metrics = [], []
    for b in range(B):
        t_img = img[b].permute(0, 3, 1, 2)                                      # [2,3,H,W]
        vec = O.extract(t_img, 21, 2)                                           # [2L,3,21,21]
        with torch.no_grad():
            params = local_m(vec).view(2, g.L, 10)
        xy, ang, eta = params[..., :4], torch.remainder(params[..., 4:8], 2 * torch.pi), params[..., 8:]
        p10 = torch.cat([xy, ang, eta], -1)
        col = O.colors_only(p10.to(F64), t_img.to(F64), g).to(F32).flatten(3, 4).flatten(1, 2).permute(0, 2, 1)    # [2,L,9]
        pm = torch.cat([xy / 3, (ang - torch.pi) / torch.pi, eta - 0.5, (col - 0.5) * 2], 2).unsqueeze(0).permute(0, 2, 1, 3).flatten(2, 3)
        with torch.no_grad():
            raw = global_m(pm)
        m = O.inference(O.restore_global(raw).to(F64), t_img.unsqueeze(0).to(F64), g, CAM, 10.39, densify)
        thres = 0.0 if densify == 'w' else 0.05
        dm = torch.where(m[5] > thres, m[4], torch.zeros_like(m[4]))
        maps.append(list(m) + [dm])
        d = dm.numpy()
        metrics.append(_eval_depth(d, gt[b:b + 1].numpy().astype(np.float64), d > 0))
    names = ('image', 'sharp', 'refoc', 'bndry', 'depth', 'conf', 'depth_map')
    tol = {'image': 3e-5, 'sharp': 1e-4, 'refoc': 3e-5, 'bndry': 3e-5, 'depth': 3e-5, 'conf': 1e-6, 'depth_map': 3e-5}   # fp32 networks + pm in between
    for k, name in enumerate(names):
        ref = torch.cat([mp[k] for mp in maps]).numpy()
        assert relmax(res[name].cpu().numpy(), ref) < tol[name], name
    np.testing.assert_allclose(res['metrics'].cpu().numpy(), np.array(metrics), rtol=0, atol=2e-4)


def test_patch_gather_equals_unfold_permute():
    from blurry_edges_b200 import Context, make_config
    S = 45
    g = geom(S)
    ctx = Context(make_config(H=S, W=S), 'cuda:0')
    img = planar_pair(synth.image_pairs(1, S, S, seed=97))[0].cuda()            # [2,3,H,W]
    vec = torch.empty(2 * g.L, 3, 21, 21, device='cuda')
    ctx.call('be_patch_gather', img, 2, vec)
    ref = torch.nn.Unfold(21, stride=2)(img).view(2, 3, 21, 21, g.Hp, g.Wp).permute(0, 4, 5, 1, 2, 3).reshape(2 * g.L, 3, 21, 21)
    assert torch.equal(vec, ref)


def test_driver_body_as_cuda_graph_equals_eager():
    """cuda_graph=True captures gather -> LocalStage -> pass A -> pm -> GlobalStage -> pass B -> metrics once and replays it."""
    import time
    from blurry_edges_b200 import DepthEstimatorFused, _lib
    S, B = GEOMS['mid'], 1
    args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], batch_size=B, mag=4.0, rho_prime=10.39,
                              densify=None, crop=10, cam_params=CAMP)
    local_m, global_m = TinyLocal().cuda(), TinyGlobal().cuda()
    eager = DepthEstimatorFused(args, local_m, global_m, 'cuda:0')
    graphed = DepthEstimatorFused(args, local_m, global_m, 'cuda:0', cuda_graph=True)
    for seed in (101, 102, 103):                       # first call captures, the others replay with new inputs
        img = synth.image_pairs(B, S, S, seed=seed).cuda()
        gt = synth.uniform((B, S, S), seed + 10, 0.75, 1.18).cuda()
        a = eager(img, gt)
        n0 = _lib.launch_count()
        b = graphed(img, gt)
        torch.cuda.synchronize()
        if seed != 101:
            assert _lib.launch_count() == n0           # a replay issues no launches from the host
        for k in a:
            assert relmax(b[k].cpu().numpy(), a[k].cpu().numpy()) < (2e-6 if k != 'metrics' else 1e-6), k
    def lat(m):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            m(img, gt)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / 20
    te, tg = lat(eager), lat(graphed)
    print(f'driver body, 1 pair {S}x{S}, stand-in networks: eager {te * 1e6:.0f} us, CUDA graph {tg * 1e6:.0f} us')
    assert tg < te


def test_cuda_graphs_are_dropped_when_the_context_is_regrown():
    """a captured graph bakes in the context's workspace pointers; a larger batch replaces the context and frees them.
    Graphs captured before must not be replayed afterwards: B=1 (capture), B=3 (context regrown, new capture), B=1 again (captured
    anew, against the live context) - every result equals the eager driver."""
    from blurry_edges_b200 import DepthEstimatorFused
    S = GEOMS['mid']
    args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], batch_size=1, mag=4.0, rho_prime=10.39,
                              densify=None, crop=10, cam_params=CAMP)
    local_m, global_m = TinyLocal().cuda(), TinyGlobal().cuda()
    eager = DepthEstimatorFused(args, local_m, global_m, 'cuda:0', max_batch=3)
    graphed = DepthEstimatorFused(args, local_m, global_m, 'cuda:0', cuda_graph=True)       # context sized for ONE pair
    for step, B in enumerate((1, 3, 1, 3)):
        img = synth.image_pairs(B, S, S, seed=120 + step).cuda()
        before = graphed.ctx
        b = {k: v.clone() for k, v in graphed(img).items()}
        if step == 1:
            assert graphed.ctx is not before and len(graphed._graphs) == 1                  # the B=1 graph went with the old context
        a = eager(img)
        torch.cuda.synchronize()
        for k in a:
            assert relmax(b[k].cpu().numpy(), a[k].cpu().numpy()) < 2e-6, (step, k)
