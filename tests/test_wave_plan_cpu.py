"""Host logic of the decoder's stage wavefront (gauspcc_b200/codec.py): the row ranges `_wave_plan` releases must never let a conv
read a row whose symbols are not decoded yet.  Brute force over real 5^3 neighbourhoods on small clouds; runs on CPU tensors."""
import types

import numpy as np
import pytest
import torch

from gauspcc_b200.codec import GausPcgcCodec


def _sorted_cloud(rng, n, ext, flat=False):
    pts = rng.integers(0, ext, size=(4 * n, 3))
    if flat:
        pts[:, 2] = rng.integers(0, 3, size=4 * n)                 # three z planes only: the plan must degenerate, not break
    pts = np.unique(pts, axis=0)[:n]
    order = np.lexsort((pts[:, 0], pts[:, 1], pts[:, 2]))         # (z, y, x) ascending == calculate_morton_order
    return pts[order]


def _keys(pts):
    p = pts.astype(np.int64) + (1 << 20)
    return (p[:, 2] << 42) | (p[:, 1] << 21) | p[:, 0]


def _max_neighbour_row(pts):
    """for every row the largest row index among its occupied 5^3 neighbours (itself included)"""
    index = {tuple(p): i for i, p in enumerate(pts.tolist())}
    out = np.arange(len(pts))
    offs = [(dx, dy, dz) for dz in range(-2, 3) for dy in range(-2, 3) for dx in range(-2, 3)]
    for i, (x, y, z) in enumerate(pts.tolist()):
        m = i
        for dx, dy, dz in offs:
            j = index.get((x + dx, y + dy, z + dz))
            if j is not None and j > m:
                m = j
        out[i] = m
    return out


@pytest.mark.parametrize("n,ext,chunk,tile,flat", [(3000, 24, 512, 64, False), (2500, 40, 256, 128, False), (1500, 12, 512, 64, False),
                                                   (2000, 30, 512, 64, True), (700, 9, 128, 32, False)])
def test_wave_plan_never_runs_ahead_of_the_decoder(n, ext, chunk, tile, flat):
    rng = np.random.default_rng(n + ext)
    pts = _sorted_cloud(rng, n, ext, flat)
    n = len(pts)
    keys = torch.from_numpy(_keys(pts))
    assert bool((keys[1:] > keys[:-1]).all())
    fake = types.SimpleNamespace(dev=torch.device("cpu"))
    chunks = [(r, min(r + chunk, n)) for r in range(0, n, chunk)]
    E, CA = GausPcgcCodec._wave_plan(fake, keys, n, chunks, tile)
    far = _max_neighbour_row(pts)
    nc = len(chunks)
    assert E[0] == [r1 for _, r1 in chunks]
    for j in range(1, 4):
        assert len(E[j]) == nc and len(CA[j]) == nc
        assert E[j][-1] == n and CA[j][-1] == n                                   # the last piece flushes the level
        assert all(a <= b for a, b in zip(E[j][:-1], E[j][1:])) and all(a <= b for a, b in zip(CA[j][:-1], CA[j][1:]))
        for c in range(nc):
            known, ca, e = E[j - 1][c], CA[j][c], E[j][c]
            assert e <= ca <= known
            assert ca == n or ca % tile == 0
            assert e == n or e % tile == 0
            # first conv of stage j over rows < ca reads stage j-1 symbols of rows < known only
            assert ca == 0 or far[:ca].max() < known
            # second conv over rows < e reads first-conv outputs of rows < ca only
            assert e == 0 or far[:e].max() < ca
    if not flat and ext >= 24:
        # the lag is a few planes per stage (here ~15 of >= 24 planes in all), not the level: something of every stage is released
        # before the last piece
        assert E[3][nc - 2] > 0
