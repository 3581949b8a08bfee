"""ShapePrefetcher (blurry_edges_b200/data.py) against the reference's own DataLoader(ShapeDataset) (data/dataset.py:6-56) when the
reference is importable, else against its per-sample semantics restated: same tuples, same layouts, same values."""
import types

import numpy as np
import pytest
import torch

import refimport
import synth
from blurry_edges_b200.data import ShapePrefetcher


def _fake(mode, n, device):
    """An object with the attributes ShapeDataset.__init__ builds (data/dataset.py:7-39), filled with hash-synthesised arrays."""
    S = 29 if mode != 'local' else 21
    d = types.SimpleNamespace(mode=mode, device=device)
    d.alpha = synth.uniform((n,), 1, 180.0, 200.0)
    if mode == 'local':
        d.img_ny, d.img_gt = synth.uniform((n, S, S, 3), 2, 0, 200), synth.uniform((n, S, S, 3), 3, 0, 200)
        d.bndry_dist, d.deri = synth.uniform((n, S, S), 4, 0, 6), synth.uniform((n, S - 2, S - 2, 3), 5, 0, 2)
    else:
        d.img_ny = synth.uniform((n, 2, S, S, 3), 2, 0, 200)
        if mode == 'global':
            d.img_gt = synth.uniform((n, 2, S, S, 3), 3, 0, 200)
            d.input_param = synth.uniform((n, 2, 25, 19), 6, -1, 1)
            d.bndry_dist, d.bndry_depth = synth.uniform((n, S, S), 4, 0, 6), synth.uniform((n, S, S), 7, 0, 1.2)
            d.deri = synth.uniform((n, 2, S - 2, S - 2, 3), 5, 0, 2)
    return d


def _getitem(d, idx):                        # data/dataset.py:40-56, restated
    a = d.alpha[idx]
    if d.mode == 'local':
        return d.img_ny[idx] / a, d.img_gt[idx] / a, d.bndry_dist[idx], d.deri[idx]
    if d.mode == 'global_pre':
        return (d.img_ny[idx] / a,)
    return d.input_param[idx], d.img_ny[idx] / a, d.img_gt[idx] / a, d.bndry_dist[idx], d.deri[idx], d.bndry_depth[idx]


def _check(device, mode, n, bs, drop_last):
    d = _fake(mode, n, device)
    batches = list(ShapePrefetcher(d, bs, device=device, drop_last=drop_last))
    assert len(batches) == (n // bs if drop_last else -(-n // bs))
    for k, got in enumerate(batches):
        got = got if isinstance(got, tuple) else (got,)
        idx = range(k * bs, min((k + 1) * bs, n))
        want = [torch.stack(f) for f in zip(*[_getitem(d, i) for i in idx])]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g.device.type == torch.device(device).type and torch.equal(g.cpu(), w)


@pytest.mark.parametrize('mode', ['global', 'local', 'global_pre'])
@pytest.mark.parametrize('drop_last', [True, False])
def test_prefetcher_yields_the_loader_batches_cpu(mode, drop_last):
    _check('cpu', mode, 11, 4, drop_last)


def test_prefetcher_shuffle_is_a_permutation():
    d = _fake('local', 10, 'cpu')
    g = torch.Generator().manual_seed(3)
    seen = torch.cat([b[2][:, 0, 0] for b in ShapePrefetcher(d, 5, device='cpu', shuffle=True, generator=g)])
    assert sorted(seen.tolist()) == sorted(d.bndry_dist[:, 0, 0].tolist())


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['global', 'local'])
def test_prefetcher_on_the_device_with_pinned_staging(mode):
    _check('cuda:0', mode, 13, 4, True)


@pytest.mark.skipif(not refimport.available(), reason='reference not mounted')
def test_prefetcher_equals_the_reference_dataloader(tmp_path):
    """The unmodified ShapeDataset + DataLoader on npy files against the prefetcher fed the same dataset object."""
    from torch.utils.data import DataLoader
    data = refimport.module('data')
    n, S = 5, 29
    rng = np.random.default_rng(0)
    for part in ('train',):
        np.save(tmp_path / f'params_src_{part}.npy', rng.normal(size=(n, 2, 25, 19)))
        np.save(tmp_path / f'images_ny_{part}.npy', rng.integers(0, 200, size=(n, 2, S, S, 3)).astype(np.float64))
        np.save(tmp_path / f'images_gt_{part}.npy', rng.uniform(0, 200, size=(n, 2, S, S, 3)))
        np.save(tmp_path / f'derivative_maps_{part}.npy', rng.uniform(0, 2, size=(n, 2, S, S, 3)))
        np.save(tmp_path / f'boundary_distances_{part}.npy', rng.integers(0, 9, size=(n, S, S)).astype(np.float64))
        np.save(tmp_path / f'boundary_depths_{part}.npy', rng.uniform(0, 1.2, size=(n, S, S)))
        np.save(tmp_path / f'alphas_{part}.npy', rng.uniform(180, 200, size=(n,)))
    ds = data.ShapeDataset('cpu', data_path=str(tmp_path), train=True, mode='global')
    ref = list(DataLoader(ds, batch_size=2, shuffle=False, drop_last=True))
    ours = list(ShapePrefetcher(ds, 2, device='cpu'))
    assert len(ref) == len(ours) == 2
    for r, o in zip(ref, ours):
        for a, b in zip(r, o):
            assert torch.equal(a, b)
