"""Synthetic tet grids and analytic SDF / mSDF fields of the shapes BASELINE.json names.

The reference never ships its grid (`data/tets/tet_grid.npz`, hmsdf.py:207-212) nor a generator that runs
without TetGen (script/get_tet_smpl.py:9-27), so the bench and the parity tests build their inputs here.
Only the *array layouts* follow the reference: `vertices` f32 (N,3) + `indices` int (F,4) for the grid file
(hmsdf.py:207-212) and `v`/`f` for the SMPL-X fitted file (script/get_tet_smpl.py:22-25).
Everything is numpy on the host; callers move the arrays to the device.
"""
from __future__ import annotations

import itertools

import numpy as np

F32 = np.float32


def kuhn_grid(res: int, dtype=np.int64):
    """res^3-class Kuhn/Freudenthal lattice on [-1,1]^3: N=(res+1)^3 vertices, F=6*res^3 tets.

    Vertex id = z*(res+1)^2 + y*(res+1) + x; each cube is split into the 6 monotone lattice paths
    0 -> p1 -> p1|p2 -> 7 (permutations of the axis bits in itertools order); tets are cube-major.
    """
    n = res + 1
    lin = np.linspace(-1.0, 1.0, n, dtype=np.float64).astype(F32)
    z, y, x = np.meshgrid(lin, lin, lin, indexing="ij")
    pos = np.stack([x, y, z], -1).reshape(-1, 3).astype(F32)
    cz, cy, cx = np.meshgrid(np.arange(res), np.arange(res), np.arange(res), indexing="ij")
    base = (cz * n * n + cy * n + cx).reshape(-1).astype(np.int64)

    def corner(bits):
        return ((bits >> 2) & 1) * n * n + ((bits >> 1) & 1) * n + (bits & 1)

    per_cube = []
    for p in itertools.permutations((1, 2, 4)):
        path = (0, p[0], p[0] | p[1], 7)
        per_cube.append(np.stack([base + corner(b) for b in path], -1))
    tets = np.stack(per_cube, 1).reshape(-1, 4)
    return pos, np.ascontiguousarray(tets.astype(dtype))


# ---------------------------------------------------------------------------------- fields
def sphere_plane_field(pos):
    """config 1/5: sdf = 0.6 - |p| (positive inside, matches occ = sdf > 0), msdf = p.y + 0.1."""
    p = pos.astype(np.float64)
    sdf = 0.6 - np.linalg.norm(p, axis=-1)
    msdf = p[:, 1] + 0.1
    return sdf.astype(F32), msdf.astype(F32)


_CAPSULES = (  # (a, b, r): torso, head, legs, arms -- "SMPL-like"
    ((0.0, -0.1, 0.0), (0.0, 0.45, 0.0), 0.18),
    ((0.0, 0.62, 0.0), (0.0, 0.70, 0.0), 0.11),
    ((-0.09, -0.15, 0.0), (-0.12, -0.85, 0.0), 0.08),
    ((0.09, -0.15, 0.0), (0.12, -0.85, 0.0), 0.08),
    ((-0.2, 0.42, 0.0), (-0.62, 0.40, 0.0), 0.055),
    ((0.2, 0.42, 0.0), (0.62, 0.40, 0.0), 0.055),
)


def capsule_sdf(pos, dilate: float = 0.0):
    p = pos.astype(np.float64)
    best = np.full(p.shape[0], -np.inf)
    for a, b, r in _CAPSULES:
        a, b = np.asarray(a), np.asarray(b)
        ab = b - a
        t = np.clip(((p - a) @ ab) / (ab @ ab), 0.0, 1.0)
        dist = np.linalg.norm(p - (a + t[:, None] * ab), axis=-1)
        best = np.maximum(best, r + dilate - dist)
    return best


def capsule_garment_field(pos):
    """config 2/3: capsule-union body SDF + T-shirt band mSDF."""
    p = pos.astype(np.float64)
    sdf = capsule_sdf(pos)
    msdf = np.minimum(p[:, 1] + 0.35, 0.50 - p[:, 1]) - 0.6 * np.maximum(np.abs(p[:, 0]) - 0.33, 0.0)
    return sdf.astype(F32), msdf.astype(F32)


def adversarial_field(pos, res: int, seed: int = 0):
    """Random fields with exact zeros (every 7th sdf, every 5th msdf) and jittered positions."""
    rng = np.random.default_rng(seed)
    n = pos.shape[0]
    sdf = rng.standard_normal(n).astype(F32)
    msdf = rng.standard_normal(n).astype(F32)
    sdf[::7] = 0
    msdf[::5] = 0
    jit = pos + (0.3 / res) * rng.standard_normal((n, 3)).astype(F32)
    return jit.astype(F32), sdf, msdf


# ---------------------------------------------------------------------------------- derived layouts
def smplx_layout_grid(res: int = 128, dilate: float = 0.1, seed: int = 0, dtype=np.int64):
    """config 3: an unstructured grid in the `script/get_tet_smpl.py` layout (dict v=(V,3) f32, f=(F,4)).

    TetGen and SMPL-X are unavailable, so: take the res^3 Kuhn lattice, keep tets whose centroid lies inside
    the capsule union dilated by `dilate`, compact the vertex ids, then scramble vertex labels and tet order
    with a fixed seed (destroys the lattice coalescing, like a real TetGen mesh).
    """
    pos, tets = kuhn_grid(res)
    cen = pos[tets].astype(np.float64).mean(1)
    keep = capsule_sdf(cen, dilate) > 0
    tets = tets[keep]
    used = np.unique(tets)
    remap = np.full(pos.shape[0], -1, np.int64)
    rng = np.random.default_rng(seed)
    remap[used] = rng.permutation(used.shape[0])
    v = np.empty((used.shape[0], 3), F32)
    v[remap[used]] = pos[used]
    f = remap[tets][rng.permutation(tets.shape[0])]
    return {"v": v, "f": np.ascontiguousarray(f.astype(dtype))}


def frame_offsets(n_verts: int, res: int, frame: int):
    """config 4: per-frame tet-vertex offsets delta_b ~ U(-1,1) * (1/res/2.1), seed = frame index
    (same scale as hmsdf.py:388 max_displacement)."""
    rng = np.random.default_rng(frame)
    return ((rng.random((n_verts, 3)) * 2.0 - 1.0) * (1.0 / res / 2.1)).astype(F32)


def surface_counts_bytes(F, N, V, Va, Fw, Fa):
    """Algorithmic bytes of one fwd+bwd extraction."""
    return 16 * F + 40 * N + 44 * Va + 28 * V + 24 * Fa + 24 * Fw
