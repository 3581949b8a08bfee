"""Big-image path (blurry_edges_test_big.py:113-190): images larger than the 147x147 network window are processed as
overlapping blocks; only the interior patch window of every block contributes to the big maps.

The reference keeps six `full_*` tensors of unfolded patches (6.7 GB at 1027x1027) and folds them at the end.  Here every
block is rendered straight into ONE interleaved accumulator [1,bigH,bigW,16] at its pixel origin (be_render_fold_blocks),
so blocks are independent work items.  A rank of a multi-GPU job renders a contiguous band of blocks into an accumulator that
covers only the image rows those blocks touch; every rank OWNS a band of image rows, one all_to_all hands each owner the other
ranks' partial sums for its rows (a few MB per rank, instead of a 67 MB whole-image reduce onto one rank or a 6.7 GB gather of
unfolded patches), the owner adds them, normalises its band, and the finished bands are gathered (or left sharded)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib
from .fused import _geometry_from_args


def block_grid(big_H, big_W, H, W, R, stride, n_margin):
    """Block stride in pixels and number of blocks per axis (blurry_edges_test_big.py:116-117)."""
    bs = (H - R + stride - stride * n_margin * 2, W - R + stride - stride * n_margin * 2)
    nb = (math.ceil((big_H - R - stride * n_margin * 2 + stride) / bs[0]),
          math.ceil((big_W - R - stride * n_margin * 2 + stride) / bs[1]))
    return bs, nb


def block_windows(big_H, big_W, H, W, R, stride, n_margin):
    """[(iv, ih, oy, ox, py0, py1, px0, px1)]: pixel origin and LOCAL patch window of every block, row-major
    (the V_s_l/V_e_l/H_s_l/H_e_l rules of blurry_edges_test_big.py:166-177)."""
    Hp, Wp = (H - R) // stride + 1, (W - R) // stride + 1
    bs, nb = block_grid(big_H, big_W, H, W, R, stride, n_margin)
    out = []
    for iv in range(nb[0]):
        for ih in range(nb[1]):
            vs, ve, hs, he = int(iv == 0), int(iv == nb[0] - 1), int(ih == 0), int(ih == nb[1] - 1)
            out.append((iv, ih, iv * bs[0], ih * bs[1], (1 - vs) * n_margin, (ve - 1) * n_margin + Hp,
                        (1 - hs) * n_margin, (he - 1) * n_margin + Wp))
    return out


def shard_blocks(nblk, rank, world):
    """Contiguous band of blocks of one rank (row-major order keeps a band's halo reads together)."""
    per, rem = divmod(nblk, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


class BigImageFused(nn.Module):
    """Pass A / pass B of every block of one big image pair and the final folds.

    args: the `get_args('eval', big=True)` namespace (img_size = block size, big_img_size, n_margin_patch, ...).
      colors(params, big_img)            params [nblk,2,L,10] -> colours [nblk,2,3,3,Hp,Wp]        (:153)
      forward(est, big_img)              est [nblk,L,12] -> six big maps + thresholded depth        (:165-190)
      render_partial / finish            the two halves of forward for block-sharded multi-GPU runs
    big_img is [2,3,bigH,bigW] (planar) on the device."""

    def __init__(self, args, depthCal=None, device='cuda:0', process_group=None):
        super().__init__()
        self.device = torch.device(device)
        self.depthCal = depthCal
        self.R, self.stride = int(args.R), int(args.stride)
        self.H, self.W = int(args.img_size[0]), int(args.img_size[1])
        self.big_H, self.big_W = int(args.big_img_size[0]), int(args.big_img_size[1])
        self.n_margin = int(args.n_margin_patch)
        self.process_group = process_group
        self.windows = block_windows(self.big_H, self.big_W, self.H, self.W, self.R, self.stride, self.n_margin)
        self.nblk = len(self.windows)
        last = self.windows[-1]
        if last[2] + self.H > self.big_H or last[3] + self.W > self.big_W:
            raise _lib.BlurryEdgesError(f'big image {self.big_H}x{self.big_W} is not {self.H}+k*{block_grid(self.big_H, self.big_W, self.H, self.W, self.R, self.stride, self.n_margin)[0][0]} (README: 147+88k)')
        self._geo = _geometry_from_args(args)
        self.ctx = _lib.Context(_lib.make_config(max_batch=self.nblk, **self._geo), self.device)
        self.H_patches, self.W_patches, self.L = self.ctx.Hp, self.ctx.Wp, self.ctx.L

    def _img(self, big_img):
        t = big_img.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(t.shape) != (2, 3, self.big_H, self.big_W):
            raise _lib.BlurryEdgesError(f'expects big_img [2,3,{self.big_H},{self.big_W}], got {tuple(t.shape)}')
        return t

    def colors(self, params, big_img):
        img = self._img(big_img)
        est = params.to(device=self.device, dtype=torch.float32).reshape(self.nblk * 2, self.L, 10).contiguous()
        items = [(m, w[2], w[3], 0, self.H_patches, 0, self.W_patches) for w in self.windows for m in (0, 1)]
        lay = _lib.single_planar_layout(self.big_H, self.big_W)
        return self.ctx.colors_blocks(est, img, lay, items).view(self.nblk, 2, 3, 3, self.H_patches, self.W_patches)

    def block_rows(self, lo, hi):
        """Image rows [a, b) that blocks [lo, hi) write: those of their interior patch windows."""
        if hi <= lo:
            return (0, 0)
        ws = self.windows[lo:hi]
        return (min(w[2] + w[4] * self.stride for w in ws), max(w[2] + (w[5] - 1) * self.stride + self.R for w in ws))

    def render_partial(self, est, big_img, lo=0, hi=None, acc=None, rows=None):
        """Render blocks [lo, hi) (est holds exactly those blocks) into an accumulator [1,rows,bigW,16] holding image rows `rows` =
        (a, b) (default: the whole image)."""
        hi = self.nblk if hi is None else hi
        a, b = (0, self.big_H) if rows is None else rows
        img = self._img(big_img)
        est = est.to(device=self.device, dtype=torch.float32).contiguous()
        if est.shape != (hi - lo, self.L, 12):
            raise _lib.BlurryEdgesError(f'expects est [{hi - lo},{self.L},12], got {tuple(est.shape)}')
        if acc is None:
            acc = torch.zeros(1, b - a, self.big_W, 16, device=self.device, dtype=torch.float32)
        blocks = [(0, w[2], w[3], w[4], w[5], w[6], w[7]) for w in self.windows[lo:hi]]
        if torch.are_deterministic_algorithms_enabled():
            msg = ('BigImageFused folds the blocks of one image with floating-point atomics and has no deterministic implementation '
                   '(the reference\'s big-image script does not ask for one)')
            if not torch.is_deterministic_algorithms_warn_only_enabled():
                raise _lib.BlurryEdgesError(msg)
            import warnings
            warnings.warn(msg)
        if blocks:
            self.ctx.render_fold_blocks(est, img, _lib.planar_layout(self.big_H, self.big_W), blocks, acc, acc_y0=a)
        return acc

    def finish(self, acc, thres=0.05, y0=0, packed=False):
        """accumulator (rows [y0, y0 + acc rows) of the image) -> (col_est [1,2,3,rows,W], col_shpd, col_refoc, bndry_est, depth,
        confidence, thresholded depth) for those rows (packed=True: plus the [16,rows,W] tensor they are views of)."""
        return tuple(self.ctx.fold_normalise(acc, thres, y0=y0, full_H=self.big_H, packed=packed))

    def forward(self, est, big_img, thres=0.05, gather=True):
        """All blocks on this device; with a process_group: est holds this rank's band of blocks (see shard_blocks).  The ranks
        exchange row bands (dist_utils.exchange_row_bands), every rank normalises the band of image rows it owns, and with
        gather=True rank 0 returns the full maps (other ranks None); gather=False returns (maps of the own band, (y0, y1)) on
        every rank - the maps stay sharded by rows, e.g. for per-rank device-to-host copies."""
        if self.process_group is None:
            return self.finish(self.render_partial(est, big_img), thres)
        import torch.distributed as dist
        from .dist_utils import exchange_row_bands, gather_row_bands, row_bands
        rank, world = dist.get_rank(self.process_group), dist.get_world_size(self.process_group)
        spans = [self.block_rows(*shard_blocks(self.nblk, r, world)) for r in range(world)]
        bands = row_bands(self.big_H, world)
        lo, hi = shard_blocks(self.nblk, rank, world)
        span = spans[rank] if hi > lo else (bands[rank][0], bands[rank][0])
        rows = (min(span[0], bands[rank][0]), max(span[1], bands[rank][1]))      # one accumulator over the rows written and the rows owned
        part = self.render_partial(est, big_img, lo, hi, rows=rows)
        band = exchange_row_bands(part[0], rows, span, spans, bands, self.process_group)
        maps = self.finish(band.unsqueeze(0), thres, y0=bands[rank][0], packed=True)
        if not gather:
            return maps[:7], bands[rank]
        full = gather_row_bands(maps[7], bands, group=self.process_group)          # [16,H,W] on rank 0
        if full is None:
            return None
        H, W = self.big_H, self.big_W
        return (full[0:6].view(1, 2, 3, H, W), full[6:9].view(1, 3, H, W), full[9:12].view(1, 3, H, W), full[12:13].view(1, 1, H, W),
                full[13:14].view(1, H, W), full[14:15].view(1, H, W), full[15:16].view(1, H, W))
