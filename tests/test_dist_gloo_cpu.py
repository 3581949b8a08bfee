"""world_size-2 gloo tests (CPU) of the two multi-GPU recipes of the path, with the oracle doing the arithmetic:
  * training: ranks hold disjoint halves of the batch, all-reduce the depth-term mask count, normalise by the global
    patch count -> the ranks' losses add up to the unmodified reference's full-batch loss (golden vector);
  * big image: ranks render contiguous bands of blocks into partial accumulators, one sum-reduce -> the stitched maps."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)


def _train_worker(rank, world, port, out):
    _init(rank, world, port)
    from blurry_edges_b200.dist_utils import sync_loss_normalisers
    from common import F64, Golden, gloss_inputs
    from oracle import be_oracle as O
    g, raw, img_ny, img_gt, bd, deri, zgt = gloss_inputs('tiny', 'normal', F64)
    gold = Golden('global_loss')
    gam = gold('tiny/gloss/normal/idx0/f64/gammas')
    sl = slice(rank, rank + 1)                                       # one sample per rank of the B=2 batch
    cam = O.Camera()
    # stage 1 on this rank: mask count of the local slice
    _, _, aux = O.global_loss(raw[sl], img_ny[sl], img_gt[sl], bd[sl], deri[sl], zgt[sl], gam, g, cam, return_terms=True)
    cnt = torch.tensor([int(aux['msum'].item())], dtype=torch.int64)
    cnt, npatch = sync_loss_normalisers(cnt, g.L, None)
    assert npatch == world * g.L
    r = raw[sl].clone().requires_grad_(True)
    loss = O.global_loss(r, img_ny[sl], img_gt[sl], bd[sl], deri[sl], zgt[sl], gam, g, cam, global_batch=world, mask_sum=float(cnt.item()))
    (grad,) = torch.autograd.grad(loss, r)
    total = loss.detach().clone()
    dist.all_reduce(total)
    ref_loss = float(gold('tiny/gloss/normal/idx0/f64/loss'))
    ref_grad = gold('tiny/gloss/normal/idx0/f64/grad')[rank:rank + 1]
    ok = abs(total.item() - ref_loss) <= 1e-10 * abs(ref_loss) and np.abs(grad.numpy() - ref_grad).max() <= 1e-8 * np.abs(ref_grad).max()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _big_worker(rank, world, port, out):
    _init(rank, world, port)
    import synth
    from blurry_edges_b200.big import block_windows, shard_blocks
    from blurry_edges_b200.dist_utils import reduce_accumulator
    from common import F64, geom, planar_pair
    from oracle import be_oracle as O
    big, g, cam = 235, geom(147), O.Camera()
    img = planar_pair(torch.from_numpy(synth.photon_pairs(1, big, big, seed=61)).to(F64) / 190.0)[0]
    est = torch.stack([O.restore_global(synth.raw_global(1, g.L, seed=70 + k, dtype=F64))[0] for k in range(4)])
    wins = block_windows(big, big, 147, 147, 21, 2, 10)
    lo, hi = shard_blocks(len(wins), rank, world)
    acc = torch.zeros(1, big, big, 4, dtype=F64)                      # planes: image0 r,g,b + boundary (enough to check the recipe)
    for k in range(lo, hi):
        iv, ih, oy, ox, py0, py1, px0, px1 = wins[k]
        r = O.inference(est[k:k + 1], img[None, :, :, oy:oy + 147, ox:ox + 147], g, cam, return_patches=True)
        P1 = r['P1'].reshape(g.Hp, g.Wp, 3, 21, 21)
        lb = r['lb'].reshape(g.Hp, g.Wp, 21, 21)
        for py in range(py0, py1):
            for px in range(px0, px1):
                y, x = oy + 2 * py, ox + 2 * px
                acc[0, y:y + 21, x:x + 21, :3] += P1[py, px].permute(1, 2, 0)
                acc[0, y:y + 21, x:x + 21, 3] += lb[py, px]
    reduce_accumulator(acc, None)
    if rank == 0:
        ref = O.inference_big(est, img, g, cam, big, big)
        n = O.cover_count(O.Geometry(H=big, W=big), F64)
        ok = torch.allclose(acc[0, :, :, :3].permute(2, 0, 1) / n, ref[0][0, 0], rtol=0, atol=1e-10) and \
            torch.allclose(acc[0, :, :, 3] / n, ref[3][0, 0], rtol=0, atol=1e-10)
        out[0] = bool(ok)
    out[rank] = out.get(rank, True)
    dist.destroy_process_group()


@pytest.mark.parametrize('worker', [_train_worker, _big_worker])
def test_two_rank_recipes(worker):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_numa_binding_helper_is_harmless_without_a_gpu():
    """bench.py calls it on every rank of a multi-rank run; without NVML / on a single-node box it must do nothing and say so."""
    import os
    from blurry_edges_b200.dist_utils import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa_node(0)
    assert info['device'] == 0 and isinstance(info['bound'], bool)
    if not info['bound']:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
