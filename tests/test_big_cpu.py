"""Host logic of the big-image path (no GPU): block windows follow blurry_edges_test_big.py:116-125,166-177 and the
rank bands partition the blocks."""
import numpy as np
import pytest

from blurry_edges_b200.big import block_grid, block_windows, shard_blocks
from oracle import be_oracle as O


@pytest.mark.parametrize('big', [235, 323, 587, 1027])
def test_block_windows_match_reference_rules(big):
    g = O.Geometry()
    bs, nb, blocks = O.big_blocks(big, big, g, 10)
    assert block_grid(big, big, 147, 147, 21, 2, 10) == (bs, nb)
    ours = block_windows(big, big, 147, 147, 21, 2, 10)
    assert len(ours) == len(blocks) == nb[0] * nb[1]
    Hp_big = (big - 21) // 2 + 1
    cover = np.zeros((Hp_big, Hp_big), int)
    for (iv, ih, oy, ox, py0, py1, px0, px1), (iv2, ih2, y0, x0, (Vs, Ve, Hs, He), (Vsl, Vel, Hsl, Hel)) in zip(ours, blocks):
        assert (iv, ih, oy, ox) == (iv2, ih2, y0, x0) and (py0, py1, px0, px1) == (Vsl, Vel, Hsl, Hel)
        # local patch (py,px) of the block is global patch (oy/2+py, ox/2+px): the windows tile the big patch grid exactly once
        assert (oy // 2 + py0, oy // 2 + py1, ox // 2 + px0, ox // 2 + px1) == (Vs, Ve, Hs, He)
        cover[Vs:Ve, Hs:He] += 1
    assert (cover == 1).all()
    assert {1027: 121, 587: 36, 323: 9, 235: 4}[big] == len(ours)


@pytest.mark.parametrize('nblk,world', [(121, 8), (121, 2), (4, 8), (9, 4), (36, 1)])
def test_shard_blocks_partition(nblk, world):
    spans = [shard_blocks(nblk, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == nblk
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1
