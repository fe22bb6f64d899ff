"""Synthetic anchor clouds with the "sparse-global / dense-local" HAC++ distribution.

HAC grows anchors on grids of 16x, 4x and 1x the voxel size (reference:
src/gs_compress/HAC/scene/gaussian_model.py:844-850, arguments/__init__.py:52-55) over scenes of
roughly +-20 units at voxel 0.001 (scripts/gs_compress/run_ours_hac.sh:7).  The generator draws
surface patches (two tangent directions with a wide Gaussian, a thin normal) and snaps samples to
those three grids, then keeps N unique voxels with signed coordinates.
"""
from __future__ import annotations

import numpy as np


def _unique_rows(p: np.ndarray) -> np.ndarray:
    key = ((p[:, 2].astype(np.int64) + (1 << 20)) << 42) | ((p[:, 1].astype(np.int64) + (1 << 20)) << 21) \
        | (p[:, 0].astype(np.int64) + (1 << 20))
    _, first = np.unique(key, return_index=True)
    return p[np.sort(first)]


def hac_like_cloud(n: int, seed: int = 0, extent_log2: int = 16) -> np.ndarray:
    """Return int32 [n,3] unique signed voxel indices (x,y,z), shuffled."""
    rng = np.random.default_rng(seed)
    ext = float(1 << extent_log2)
    n_patch = max(8, n // 4000)
    centres = rng.uniform(0.2, 0.8, size=(n_patch, 3)) * ext
    # random orthonormal frames
    q, _ = np.linalg.qr(rng.normal(size=(n_patch, 3, 3)))
    radius = rng.uniform(200.0, 1500.0, size=n_patch) * (ext / 65536.0)
    grids = np.array([16, 4, 1], dtype=np.float64)
    chunks = []
    have = 0
    pts = np.zeros((0, 3), dtype=np.int32)
    per = 4000
    while have < n:
        todo = max(8, int((n - have) * 1.35 / per) + 1)
        pid = rng.integers(0, n_patch, size=todo)
        t = rng.normal(size=(todo, per, 2)) * radius[pid][:, None, None]
        nrm = rng.normal(size=(todo, per, 1)) * 3.0
        local = np.concatenate([t, nrm], axis=-1)                       # [todo, per, 3]
        world = centres[pid][:, None, :] + np.einsum("bpk,bkj->bpj", local, q[pid].transpose(0, 2, 1))
        g = grids[rng.choice(3, size=(todo, per, 1), p=[0.25, 0.35, 0.40])]
        snapped = np.round(world / g) * g
        snapped = np.clip(snapped, 0, ext - 1).reshape(-1, 3).astype(np.int32)
        chunks.append(snapped)
        pts = _unique_rows(np.concatenate([pts] + chunks, axis=0))
        chunks = []
        have = pts.shape[0]
    perm = rng.permutation(pts.shape[0])[:n]
    out = pts[perm] - np.int32(1 << (extent_log2 - 1))
    return np.ascontiguousarray(out.astype(np.int32))


def uniform_unique_cloud(n: int, seed: int = 0, extent_log2: int = 17) -> np.ndarray:
    """Uniform unique voxels in a 2^extent cube (kernel-map / ordering sweeps, BASELINE config 3)."""
    rng = np.random.default_rng(seed)
    ext = 1 << extent_log2
    pts = np.zeros((0, 3), dtype=np.int32)
    while pts.shape[0] < n:
        add = rng.integers(0, ext, size=(int((n - pts.shape[0]) * 1.1) + 16, 3), dtype=np.int32)
        pts = _unique_rows(np.concatenate([pts, add], axis=0))
    perm = rng.permutation(pts.shape[0])[:n]
    return np.ascontiguousarray(pts[perm] - np.int32(ext // 2))
