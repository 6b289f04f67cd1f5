"""CPU oracle for the triangle-mesh stage behind the extraction.

TEST INFRASTRUCTURE ONLY (same rules as oracle/gshell_oracle.py: imported by tests/, smoke() and the CPU legs of the
bench scripts, never by the product package).  Numpy restatement of

    Mesh.get_edge    render/mesh.py:240-250   3F (min,max) rows -> torch.unique(dim=0)
    auto_normals     render/mesh.py:418-446   face normals scatter-added to the vertices, degenerate -> (0,0,1), normalised

Parity pin: checked against the live reference module (loaded by oracle/ref_loader.load_reference_mesh) in
tests/test_mesh_oracle.py and against golden vectors produced from it by oracle/make_golden_mesh.py
(tests/golden/mesh_*.npz).  Edges are bit-exact; normals follow the CPU reference's op order (sequential
scatter: all first corners, then all second, then all third) and agree to 1e-6 absolute (a few ulp: ATen's
vectorised sqrt / divide); the gradient is the hand-derived adjoint evaluated in float64 (1e-5 relative against the
reference's autograd).
"""
from __future__ import annotations

import numpy as np

from .gshell_oracle import F32, _cross_f32, _finish_normals


def mesh_edges(faces: np.ndarray) -> np.ndarray:
    """render/mesh.py:240-250 -> (E,2) int64, rows ascending lexicographically."""
    faces = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    if faces.shape[0] == 0:
        return np.zeros((0, 2), np.int64)
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    e = np.sort(e, axis=1)
    key = (e[:, 0].astype(np.uint64) << np.uint64(32)) | e[:, 1].astype(np.uint64)
    key = np.unique(key)
    return np.stack([(key >> np.uint64(32)).astype(np.int64), (key & np.uint64(0xFFFFFFFF)).astype(np.int64)], 1)


def _face_sides(pos, faces):
    v0, v1, v2 = pos[faces[:, 0]], pos[faces[:, 1]], pos[faces[:, 2]]
    return (v1 - v0).astype(F32), (v2 - v0).astype(F32)


def normal_sums(pos: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """render/mesh.py:420-437: un-normalised vertex normal sums (V,3) fp32."""
    pos = np.asarray(pos, F32)
    faces = np.asarray(faces, np.int64).reshape(-1, 3)
    acc = np.zeros((pos.shape[0], 3), F32)
    if faces.shape[0] == 0:
        return acc
    a, b = _face_sides(pos, faces)
    # torch.cross without `dim` takes the FIRST axis of size 3: the face axis when there are exactly three faces
    fn = _cross_f32(a, b, 0 if faces.shape[0] == 3 else -1)
    for c in range(3):
        np.add.at(acc, faces[:, c], fn)
    return acc


def normal_condition(pos: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """Per vertex: sum of |face normal| over |sum of face normals| (1 for vertices without faces; corners of faces with a
    repeated index count the face's squared side instead of its vanishing normal, see below).  The fp32 sum is
    accumulated in a different order by float atomics (the reference's CUDA scatter_add included), which moves the unit
    normal by about 1e-7 x this number; vertices whose normals cancel (1e4 and above) have no stable normal at all."""
    pos = np.asarray(pos, F32)
    faces = np.asarray(faces, np.int64).reshape(-1, 3)
    cond = np.ones(pos.shape[0])
    if faces.shape[0] == 0:
        return cond
    a, b = _face_sides(pos, faces)
    fn = _cross_f32(a, b, 0 if faces.shape[0] == 3 else -1).astype(np.float64)
    mag = np.zeros(pos.shape[0])
    for c in range(3):
        np.add.at(mag, faces[:, c], np.linalg.norm(fn, axis=1))
    s = np.linalg.norm(normal_sums(pos, faces).astype(np.float64), axis=1)
    used = mag > 0
    if not used.any():  # no face has a non-zero normal: every vertex keeps the default (0,0,1)
        return cond
    # A face with a repeated index has a zero normal in exact arithmetic, but the fused cross product leaves a rounding
    # residue (~1e-9).  A vertex that only belongs to such faces gets that residue as its "sum", a 1 / |sum| ~ 1e9 gradient,
    # and -- because the gradient of the repeated corner is added and subtracted at the same vertex -- swamps the fp32
    # gradients of its neighbours (in the reference as well).  For the corners of such faces the sum is therefore counted
    # against the face's side length, not against its (vanishing) normal.
    rep = (faces[:, 0] == faces[:, 1]) | (faces[:, 1] == faces[:, 2]) | (faces[:, 2] == faces[:, 0])
    if rep.any():
        side2 = np.maximum((a.astype(np.float64) ** 2).sum(-1), (b.astype(np.float64) ** 2).sum(-1))
        for c in range(3):
            np.maximum.at(mag, faces[rep, c], 1e-2 * side2[rep])
    cond[used] = mag[used] / np.maximum(s[used], 1e-300)
    return cond


def auto_normals(pos: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """render/mesh.py:418-441 -> v_nrm (V,3) fp32."""
    return _finish_normals(normal_sums(pos, faces))


def auto_normals_backward(pos: np.ndarray, faces: np.ndarray, g_nrm: np.ndarray) -> np.ndarray:
    """d loss / d pos for upstream g_nrm (V,3): adjoint of auto_normals, float64 arithmetic, returned as fp32."""
    pos64 = np.asarray(pos, np.float64)
    faces = np.asarray(faces, np.int64).reshape(-1, 3)
    g = np.asarray(g_nrm, np.float64)
    g_pos = np.zeros_like(pos64)
    if faces.shape[0] == 0:
        return g_pos.astype(F32)
    acc32 = normal_sums(pos, faces)
    d32 = np.sum((acc32 * acc32).astype(F32), -1, dtype=F32)
    keep = d32 > F32(1e-20)  # the forward's choice, taken in fp32
    s = acc32.astype(np.float64)
    ln = np.sqrt(np.maximum((s * s).sum(-1), 1e-300))
    n = s / ln[:, None]
    gs = (g - n * (n * g).sum(-1, keepdims=True)) / ln[:, None]
    gs[~keep] = 0.0
    G = gs[faces[:, 0]] + gs[faces[:, 1]] + gs[faces[:, 2]]  # (F,3) d loss / d face normal
    a = pos64[faces[:, 1]] - pos64[faces[:, 0]]
    b = pos64[faces[:, 2]] - pos64[faces[:, 0]]
    if faces.shape[0] == 3:  # column c of the face-normal matrix is a[:,c] x b[:,c]
        ga = np.cross(b, G, axis=0)
        gb = np.cross(G, a, axis=0)
    else:
        ga = np.cross(b, G)
        gb = np.cross(G, a)
    np.add.at(g_pos, faces[:, 1], ga)
    np.add.at(g_pos, faces[:, 2], gb)
    np.add.at(g_pos, faces[:, 0], -(ga + gb))
    return g_pos.astype(F32)
