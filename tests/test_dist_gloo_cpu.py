"""world_size-2 gloo tests (CPU) of the two multi-GPU recipes of the path, with the oracle doing the arithmetic:
  * training: ranks hold disjoint halves of the batch, all-reduce the depth-term mask count, normalise by the global
    patch count -> the ranks' losses add up to the unmodified reference's full-batch loss (golden vector);
  * big image: ranks render contiguous bands of blocks into partial accumulators over the rows they touch, one all_to_all hands
    every rank the partial sums of the row band it owns -> that band of the stitched maps."""
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
    from blurry_edges_b200.dist_utils import exchange_row_bands, row_bands
    from common import F64, geom, planar_pair
    from oracle import be_oracle as O
    big, g, cam = 235, geom(147), O.Camera()
    img = planar_pair(torch.from_numpy(synth.photon_pairs(1, big, big, seed=61)).to(F64) / 190.0)[0]
    est = torch.stack([O.restore_global(synth.raw_global(1, g.L, seed=70 + k, dtype=F64))[0] for k in range(4)])
    wins = block_windows(big, big, 147, 147, 21, 2, 10)
    # the row spans every rank's blocks write and the row bands the ranks own (BigImageFused.forward does the same on the device)
    def rows_of(lo_, hi_):
        ws = wins[lo_:hi_]
        return (min(w[2] + 2 * w[4] for w in ws), max(w[2] + 2 * (w[5] - 1) + 21 for w in ws)) if ws else (0, 0)
    spans = [rows_of(*shard_blocks(len(wins), r, world)) for r in range(world)]
    bands = row_bands(big, world)
    lo, hi = shard_blocks(len(wins), rank, world)
    rows = (min(spans[rank][0], bands[rank][0]), max(spans[rank][1], bands[rank][1]))      # rows written + rows owned
    a = rows[0]
    acc = torch.zeros(rows[1] - rows[0], big, 4, dtype=F64)           # planes: image0 r,g,b + boundary (enough to check the recipe)
    for k in range(lo, hi):
        iv, ih, oy, ox, py0, py1, px0, px1 = wins[k]
        r = O.inference(est[k:k + 1], img[None, :, :, oy:oy + 147, ox:ox + 147], g, cam, return_patches=True)
        P1 = r['P1'].reshape(g.Hp, g.Wp, 3, 21, 21)
        lb = r['lb'].reshape(g.Hp, g.Wp, 21, 21)
        for py in range(py0, py1):
            for px in range(px0, px1):
                y, x = oy + 2 * py - a, ox + 2 * px
                acc[y:y + 21, x:x + 21, :3] += P1[py, px].permute(1, 2, 0)
                acc[y:y + 21, x:x + 21, 3] += lb[py, px]
    band = exchange_row_bands(acc, rows, spans[rank], spans, bands, None)   # ONE all_to_all: every owner gets the partial sums of its rows
    y0, y1 = bands[rank]
    ref = O.inference_big(est, img, g, cam, big, big)
    n = O.cover_count(O.Geometry(H=big, W=big), F64)[y0:y1]
    ok = torch.allclose(band[:, :, :3].permute(2, 0, 1) / n, ref[0][0, 0][:, y0:y1], rtol=0, atol=1e-10) and \
        torch.allclose(band[:, :, 3] / n, ref[3][0, 0][y0:y1], rtol=0, atol=1e-10)
    # the finished bands go to rank 0 in one gather (one message per rank)
    from blurry_edges_b200.dist_utils import gather_row_bands
    mine = torch.cat([band[:, :, :3].permute(2, 0, 1) / n, (band[:, :, 3] / n).unsqueeze(0)]).contiguous()      # [4 planes, band rows, W]
    full = gather_row_bands(mine, bands)
    if rank == 0:
        ok = ok and torch.allclose(full[:3], ref[0][0, 0], rtol=0, atol=1e-10) and torch.allclose(full[3], ref[3][0, 0], rtol=0, atol=1e-10)
    else:
        ok = ok and full is None
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize('worker,world', [(_train_worker, 2), (_big_worker, 2), (_big_worker, 3)])
def test_multi_rank_recipes(worker, world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}


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
