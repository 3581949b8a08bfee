"""Two-rank NCCL tests of the two multi-GPU recipes of the path on real devices (skipped on a box with one GPU; the CPU
suite runs the same recipes under gloo with the oracle doing the arithmetic, tests/test_dist_gloo_cpu.py):
  * training: each rank holds one sample of the golden B=2 batch, GlobalLossFused(process_group=...) all-reduces the mask
    count between its two kernel stages -> the ranks' losses add up to the full-batch loss of the unmodified reference and
    every rank's gradient is its slice of the full-batch gradient;
  * big image: each rank renders its band of blocks (shard_blocks), BigImageFused sums the partial accumulators onto rank 0
    -> the maps of the single-GPU run."""
import argparse
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
CAMP = {'s': 0.1104, 'rho_1': 10.0, 'rho_2': 10.2, 'sigma_cam': 0.003, 'pixel_pitch': 5.86e-6}


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    import sys
    for p in (ROOT, os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))


def _train_worker(rank, world, port, out):
    _init(rank, world, port)
    from blurry_edges_b200 import GlobalLossFused
    from common import F32, GEOMS, Golden, gloss_inputs
    dev = f'cuda:{rank}'
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('tiny', 'normal', F32)
    gold = Golden('global_loss')
    gam = gold('tiny/gloss/normal/idx0/f64/gammas')
    args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[GEOMS['tiny']] * 2, batch_size=1, mag=4.0, cam_params=CAMP,
                              gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                              gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                              gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])
    crit = GlobalLossFused(args, None, dev, process_group=dist.group.WORLD, grad_reduce='sum')    # shares that ADD UP
    crit.update_gamma()
    np.testing.assert_allclose(crit.gammas(), gam, rtol=0, atol=0)
    sl = slice(rank, rank + 1)                                       # one sample per rank of the B=2 batch
    est = raw[sl].clone().to(dev).requires_grad_(True)
    loss = crit(est, img_ny[sl].to(dev), img_gt[sl].to(dev), bd[sl].to(dev), deri[sl].to(dev), zgt[sl].to(dev))
    loss.backward()
    total = loss.detach().clone()
    dist.all_reduce(total)
    ref_loss = float(gold('tiny/gloss/normal/idx0/f64/loss'))
    ref_grad = gold('tiny/gloss/normal/idx0/f64/grad')[rank:rank + 1]
    e_loss = abs(total.item() - ref_loss) / abs(ref_loss)
    e_grad = float(np.abs(est.grad.cpu().numpy() - ref_grad).max() / np.abs(ref_grad).max())
    out[rank] = (e_loss < 5e-6 and e_grad < 5e-5, e_loss, e_grad)
    dist.destroy_process_group()


def _big_worker(rank, world, port, out):
    _init(rank, world, port)
    import synth
    from blurry_edges_b200 import BigImageFused, shard_blocks
    from common import geom, planar_pair
    from oracle import be_oracle as O
    dev = f'cuda:{rank}'
    big, g = 323, geom(147)
    args = argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[147, 147], big_img_size=[big, big], batch_size=1,
                              mag=4.0, rho_prime=10.39, densify=None, n_margin_patch=10, cam_params=CAMP)
    img = planar_pair(torch.from_numpy(synth.photon_pairs(1, big, big, seed=63)).float() / 190.0)[0].to(dev)
    est = torch.stack([O.restore_global(synth.raw_global(1, g.L, seed=170 + k))[0] for k in range(9)]).to(dev)
    sharded = BigImageFused(args, None, dev, process_group=dist.group.WORLD)
    lo, hi = shard_blocks(sharded.nblk, rank, world)
    maps = sharded(est[lo:hi], img)
    if rank == 0:
        single = BigImageFused(args, None, dev)(est, img)          # all nine blocks on this GPU
        worst = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(maps, single))
        out[0] = (worst < 1e-6, worst)                              # same sums in a different order
    else:
        out[rank] = (maps is None, 0.0)
    dist.destroy_process_group()


def _gargs(S, B):
    return argparse.Namespace(R=21, stride=2, w=1.0, alpha_lambda=5e-3, img_size=[S, S], batch_size=B, mag=4.0, cam_params=CAMP,
                              gamma_color=[1.0, 0.1, 0.1], gamma_color_cons=[0.2, 0.1, 0.05], gamma_bndry_cons=[0.05, 0.05, 0.02],
                              gamma_smthns=[0.005, 0.1, 0.002], gamma_smthns_cons=[0.005, 0.1, 0.002], gamma_bndry_loc=[0.0001, 0.05, 0.0001],
                              gamma_depth=[0.0001, 0.05, 0.5], dynamic_epoch=[30, 100, 200])


def _uneven_worker(rank, world, port, out):
    """A last batch without drop_last: rank 0 holds 1 pair, rank 1 holds 2 pairs of a 3-pair batch.  The shares must add up to
    the single-process loss of the 3-pair batch and every rank's gradient must be its slice of the full gradient: the 16-byte
    all-reduce carries (mask count, patch count) (global_training.py:127 and the means of :130-139 run over the WHOLE batch)."""
    _init(rank, world, port)
    from blurry_edges_b200 import GlobalLossFused
    from common import F32, GEOMS, gloss_inputs
    dev = f'cuda:{rank}'
    S = GEOMS['mid']
    g, raw, img_ny, img_gt, bd, deri, zgt = [t.to(dev) if torch.is_tensor(t) else t for t in gloss_inputs('mid', 'normal', F32, B=3)]
    full = GlobalLossFused(_gargs(S, 3), None, dev)
    full.update_gamma()
    rf = raw.clone().requires_grad_(True)
    lf = full(rf, img_ny, img_gt, bd, deri, zgt)
    lf.backward()
    sl = slice(0, 1) if rank == 0 else slice(1, 3)
    crit = GlobalLossFused(_gargs(S, 2), None, dev, process_group=dist.group.WORLD, grad_reduce='sum')
    crit.update_gamma()
    rs = raw[sl].clone().requires_grad_(True)
    ls = crit(rs, img_ny[sl], img_gt[sl], bd[sl], deri[sl], zgt[sl])
    ls.backward()
    tot = ls.detach().clone()
    dist.all_reduce(tot)
    e_loss = float((tot - lf.detach()).abs() / lf.detach().abs())
    e_grad = float((rs.grad - rf.grad[sl]).abs().max() / rf.grad.abs().max())
    out[rank] = (e_loss < 2e-6 and e_grad < 2e-6, e_loss, e_grad)
    dist.destroy_process_group()


def _ddp_worker(rank, world, port, out):
    """stock DistributedDataParallel AVERAGES parameter gradients.  With the default grad_reduce='mean' the
    averaged gradients of a stand-in network equal the gradients of a single process holding the whole batch, and the mean of the
    ranks' losses is the whole-batch loss - so clip_grad_norm_ / AdamW see the reference's step (global_training.py:208-213)."""
    _init(rank, world, port)
    from torch.nn.parallel import DistributedDataParallel as DDP
    from blurry_edges_b200 import GlobalLossFused
    from common import F32, GEOMS, gloss_inputs
    dev = f'cuda:{rank}'
    S = GEOMS['tiny']
    g, raw, img_ny, img_gt, bd, deri, zgt = [t.to(dev) if torch.is_tensor(t) else t for t in gloss_inputs('tiny', 'normal', F32, B=2)]

    def make_net():
        torch.manual_seed(5)
        return torch.nn.Linear(12, 12).to(dev)                    # stand-in for GlobalStage: est = raw + 0.05 * net(raw)

    net1 = make_net()
    full = GlobalLossFused(_gargs(S, 2), None, dev)
    full.update_gamma()
    l_full = full(raw + 0.05 * net1(raw), img_ny, img_gt, bd, deri, zgt)
    l_full.backward()
    net2 = DDP(make_net(), device_ids=[rank])
    crit = GlobalLossFused(_gargs(S, 1), None, dev, process_group=dist.group.WORLD)          # grad_reduce='mean' is the default
    crit.update_gamma()
    sl = slice(rank, rank + 1)
    l_rank = crit(raw[sl] + 0.05 * net2(raw[sl]), img_ny[sl], img_gt[sl], bd[sl], deri[sl], zgt[sl])
    l_rank.backward()
    mean_loss = l_rank.detach().clone()
    dist.all_reduce(mean_loss)
    mean_loss /= world
    e_loss = float((mean_loss - l_full.detach()).abs() / l_full.detach().abs())
    e_w = float((net2.module.weight.grad - net1.weight.grad).abs().max() / net1.weight.grad.abs().max())
    e_b = float((net2.module.bias.grad - net1.bias.grad).abs().max() / net1.bias.grad.abs().max())
    out[rank] = (e_loss < 2e-6 and e_w < 1e-5 and e_b < 1e-5, e_loss, e_w, e_b)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
@pytest.mark.parametrize('worker', [_train_worker, _big_worker, _uneven_worker, _ddp_worker])
def test_two_rank_nccl_recipes(worker):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = dict(out)
    assert set(res) == {0, 1} and all(v[0] for v in res.values()), res
