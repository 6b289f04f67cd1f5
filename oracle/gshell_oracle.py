"""CPU oracle for the G-Shell / mSDF marching-tetrahedra extraction path.

TEST INFRASTRUCTURE ONLY.  This file is a numpy restatement of the algorithm the
reference implements with PyTorch ops in

    geometry/gshell_tets.py:253-447        (GShell_Tets.__call__)
    geometry/hmsdf_tets_split.py:254-454   (hmSDF_Tets.__call__, adds the `type` sign flip :261-264)

It exists so that the CUDA path can be checked on a machine where the reference
tree is not present.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product
package never does (and fails loudly without its CUDA library).

Parity pin: the oracle is checked bit-for-bit (integers, positions, mSDF values)
and to 1e-6 / 1e-5 (tangents / gradients) against
  * the live reference, when `/root/reference` is present (tests/test_oracle_vs_reference.py), and
  * golden vectors produced from the live reference by `oracle/make_golden.py`
    and committed under `tests/golden/` (tests/test_oracle_golden.py).

Structure differs from the reference on purpose: edges are de-duplicated by sorting a
64-bit (min,max) key and run-length encoding (the reference calls torch.unique(dim=0)),
polygon corners are kept in one flat "corner array", and the backward pass is the
hand-derived adjoint (the reference relies on autograd).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# --------------------------------------------------------------------------------------
# Case tables (data; element-for-element equal to gshell_tets.py:91-190, asserted by tests)
# --------------------------------------------------------------------------------------
_X = -1
#: marching-tets triangles per 4-bit occupancy code, as tet-edge ids (gshell_tets.py:91-108)
TRIANGLE_TABLE = np.array([
    (_X,) * 6, (1, 0, 2, _X, _X, _X), (4, 0, 3, _X, _X, _X), (1, 4, 2, 1, 3, 4),
    (3, 1, 5, _X, _X, _X), (2, 3, 0, 2, 5, 3), (1, 4, 0, 1, 5, 4), (4, 2, 5, _X, _X, _X),
    (4, 5, 2, _X, _X, _X), (4, 1, 0, 4, 5, 1), (3, 2, 0, 3, 5, 2), (1, 3, 5, _X, _X, _X),
    (4, 1, 2, 4, 3, 1), (3, 0, 4, _X, _X, _X), (2, 0, 1, _X, _X, _X), (_X,) * 6,
], dtype=np.int64)
#: polygon loop (closed) per occupancy code, as tet-edge ids (gshell_tets.py:110-127)
MESH_EDGE_TABLE = np.array([
    (_X,) * 6, (1, 0, 2, 1, _X, _X), (4, 0, 3, 4, _X, _X), (1, 3, 4, 2, 1, _X),
    (3, 1, 5, 3, _X, _X), (2, 5, 3, 0, 2, _X), (1, 5, 4, 0, 1, _X), (4, 2, 5, 4, _X, _X),
    (4, 5, 2, 4, _X, _X), (4, 5, 1, 0, 4, _X), (3, 5, 2, 0, 3, _X), (1, 3, 5, 1, _X, _X),
    (4, 3, 1, 2, 4, _X), (3, 0, 4, 3, _X, _X), (2, 0, 1, 2, _X, _X), (_X,) * 6,
], dtype=np.int64)
#: mSDF cut of a triangle polygon; locals 0-2 corners, 3-5 boundary verts (gshell_tets.py:130-147)
TRIANGLE_TABLE_TRI = np.array([
    (_X,) * 6, (4, 2, 5, _X, _X, _X), (3, 1, 4, _X, _X, _X), (3, 1, 2, 3, 2, 5),
    (0, 3, 5, _X, _X, _X), (0, 3, 4, 0, 4, 2), (0, 1, 4, 0, 4, 5), (0, 1, 2, _X, _X, _X),
], dtype=np.int64)
#: mSDF cut of a quad polygon; locals 0-3 corners, 4-7 boundary verts (gshell_tets.py:149-184)
TRIANGLE_TABLE_QUAD = np.array([
    (_X,) * 12,
    (6, 3, 7) + (_X,) * 9,
    (5, 2, 6) + (_X,) * 9,
    (5, 2, 7, 3, 7, 2) + (_X,) * 6,
    (4, 1, 5) + (_X,) * 9,
    (4, 1, 5, 4, 5, 7, 5, 6, 7, 7, 6, 3),
    (4, 1, 2, 6, 4, 2) + (_X,) * 6,
    (4, 1, 2, 7, 4, 2, 7, 2, 3) + (_X,) * 3,
    (0, 4, 7) + (_X,) * 9,
    (0, 4, 6, 3, 0, 6) + (_X,) * 6,
    (0, 4, 5, 0, 5, 2, 0, 2, 6, 0, 6, 7),
    (0, 4, 5, 0, 5, 2, 0, 2, 3) + (_X,) * 3,
    (0, 1, 5, 7, 0, 5) + (_X,) * 6,
    (0, 1, 5, 0, 5, 6, 0, 6, 3) + (_X,) * 3,
    (0, 1, 2, 0, 2, 6, 0, 6, 7) + (_X,) * 3,
    (0, 1, 2, 0, 2, 3) + (_X,) * 6,
], dtype=np.int64)
NUM_TRIANGLES_TABLE = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0], dtype=np.int64)  # :186
NUM_TRIANGLES_TRI_TABLE = np.array([0, 1, 1, 2, 1, 2, 2, 1], dtype=np.int64)  # :189
NUM_TRIANGLES_QUAD_TABLE = np.array([0, 1, 1, 2, 1, 4, 2, 3, 1, 2, 4, 3, 2, 3, 3, 2], dtype=np.int64)  # :190
#: endpoints of the 6 tet edges, flattened (gshell_tets.py:187)
BASE_TET_EDGES = np.array([0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3], dtype=np.int64)
_EDGE_P = BASE_TET_EDGES[0::2]
_EDGE_Q = BASE_TET_EDGES[1::2]

EPS12 = F32(1e-12)


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def linspace_f32(n: int) -> np.ndarray:
    """torch.linspace(0, 1 - 1/n, n, dtype=float32) as evaluated on CPU (gshell_tets.py:221-224).

    ATen fills the lower half as start + step*i and the upper half as end - step*(n-1-i),
    the latter with one rounding (vectorised fmadd).  Verified element-exact for n in
    {1..17, 100, 1254, 3548, 10033, 24577} in tests/test_oracle_vs_reference.py.
    """
    end = F32(1 - (1 / n))
    step = end / F32(n - 1) if n > 1 else F32(0)
    i = np.arange(n, dtype=np.int64)
    lo = (step * i.astype(F32)).astype(F32)
    hi = (np.float64(end) - np.float64(step) * (n - 1 - i)).astype(F32)  # exact product, single rounding
    return np.where(i < n // 2, lo, hi).astype(F32)


def vertex_uv(k: np.ndarray, num_tets: int) -> np.ndarray:
    """UV the reference ends up using for *vertex id* k (quirk: gshell_tets.py:327 passes `faces`
    as the texture index, so the per-tet UV atlas of map_uv (:219-233) is indexed by vertex ids)."""
    nuv = int(np.ceil(np.sqrt((num_tets * 2 + 1) // 2)))
    lin = linspace_f32(nuv)
    pad = F32(0.9 / nuv)
    cell, c = k >> 2, k & 3
    ix, iy = cell % nuv, cell // nuv
    u = lin[ix]
    v = lin[iy]
    u = np.where((c == 1) | (c == 2), u + pad, u).astype(F32)
    v = np.where((c == 2) | (c == 3), v + pad, v).astype(F32)
    return np.stack([u, v], -1)


def _lerp2(xa, xb, wa, wb):
    """fl(xa*wa) + fl(xb*wb) in float32, two roundings for the products and one for the sum."""
    return ((xa * wa).astype(F32) + (xb * wb).astype(F32)).astype(F32)


def _fma_f32(x, y, z):
    """round_f32(x*y + z): the f32 product is exact in f64, so this is a single-rounding fused multiply-add."""
    return (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(F32)


def _cross_f32(a, b, axis):
    """torch.cross on CPU evaluates each component as fma(a_i, b_j, -fl(a_j*b_i)) (ATen's scalar loop is
    compiled with fp-contraction); matters only for degenerate faces, where it leaves a rounding residue
    instead of an exact zero.  Checked element-exact against torch.cross in tests/test_oracle_vs_reference.py."""
    a = np.moveaxis(a, axis, 0)
    b = np.moveaxis(b, axis, 0)
    comps = [_fma_f32(a[i], b[j], -(a[j] * b[i]).astype(F32)) for i, j in ((1, 2), (2, 0), (0, 1))]
    return np.moveaxis(np.stack(comps, 0), 0, axis).astype(F32)


def _safe_normalize(x):
    """render/util.py:25-29"""
    d = np.sum((x * x).astype(F32), -1, keepdims=True, dtype=F32)
    return (x / np.sqrt(np.maximum(d, F32(1e-20)))).astype(F32)


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def _classify(occ, mpos, tets, n_threads):
    """4-bit occupancy code and validity of every tet (gshell_tets.py:260-275, 307-308); the only O(F) stage.
    Chunked over a thread pool when n_threads > 1 (numpy releases the GIL in take / compare)."""
    def one(chunk):
        code = (occ[chunk] * np.array([1, 2, 4, 8])).sum(-1)
        valid = (code != 0) & (code != 15)
        if mpos is not None:
            valid &= mpos[chunk].any(-1)
        ids = np.nonzero(valid)[0]
        return ids, code[ids]
    n = tets.shape[0]
    if n_threads <= 1 or n < (1 << 16):
        return one(tets)
    from concurrent.futures import ThreadPoolExecutor
    bounds = np.linspace(0, n, 4 * n_threads + 1).astype(np.int64)
    with ThreadPoolExecutor(n_threads) as pool:
        parts = list(pool.map(lambda i: one(tets[bounds[i]:bounds[i + 1]]), range(4 * n_threads)))
    ids = np.concatenate([p[0] + bounds[i] for i, p in enumerate(parts)])
    return ids, np.concatenate([p[1] for p in parts])


def extract_forward(pos, sdf, msdf, tets, msdf_sign: int = 1, output_watertight_template: bool = True,
                    n_threads: int = 1):
    """Forward pass.  Returns a dict holding every returned tensor of the reference plus the
    integer intermediates the parity tests compare (valid ids, codes, sorted edge keys, corner array).

    pos (N,3) f32, sdf (N,) or (N,1) any float, msdf (N,) f32, tets (F,4) int.
    msdf_sign = -1 reproduces hmSDF_Tets(type="body") (hmsdf_tets_split.py:261-264).
    """
    pos = np.ascontiguousarray(pos, dtype=F32)
    s = np.ascontiguousarray(sdf).reshape(-1).astype(F32)  # .float(), gshell_tets.py:254
    m = np.ascontiguousarray(msdf, dtype=F32).reshape(-1)
    if msdf_sign < 0:
        m = -m
    tets = np.ascontiguousarray(tets)
    if tets.dtype != np.int64:
        tets = tets.astype(np.int64)
    n_grid = pos.shape[0]
    num_tets = tets.shape[0]

    # --- classification (gshell_tets.py:260-275, 307-309) ---
    occ = s > 0
    valid_ids, code = _classify(occ, None if output_watertight_template else (m > 0), tets, n_threads)
    tv = tets[valid_ids]
    ntri = NUM_TRIANGLES_TABLE[code]
    fv = valid_ids.shape[0]

    # --- edges: (min,max) keys, sort + run-length (gshell_tets.py:277-287) ---
    p, q = tv[:, _EDGE_P], tv[:, _EDGE_Q]  # (Fv,6)
    lo, hi = np.minimum(p, q), np.maximum(p, q)
    key = (lo << 32) | hi
    flat = key.reshape(-1)
    order = np.argsort(flat, kind="stable")
    skey = flat[order]
    head = np.ones(skey.shape[0], dtype=bool)
    head[1:] = skey[1:] != skey[:-1]
    ukey = skey[head]
    rank_sorted = np.cumsum(head) - 1
    inverse = np.empty_like(rank_sorted)
    inverse[order] = rank_sorted
    ua, ub = ukey >> 32, ukey & 0xFFFFFFFF
    crossing = occ[ua] != occ[ub] if ukey.size else np.zeros(0, bool)
    vid_of_unique = np.where(crossing, np.cumsum(crossing) - 1, -1)
    idx6 = vid_of_unique[inverse].reshape(-1, 6)
    ea, eb = ua[crossing], ub[crossing]  # interp_v, a < b
    nv = ea.shape[0]

    # --- zero-crossing interpolation (gshell_tets.py:291-303), SURVEY A.4 ---
    e0 = s[ea]
    e1 = (-s[eb]).astype(F32)
    d = (e0 + e1).astype(F32)
    dd = (np.sign(d) * (np.abs(d) + EPS12)).astype(F32)
    dd = np.where(dd == 0, EPS12, dd).astype(F32)
    w0 = (e1 / dd).astype(F32)
    w1 = (e0 / dd).astype(F32)
    verts = _lerp2(pos[ea], pos[eb], w0[:, None], w1[:, None])
    msdf_vert = _lerp2(m[ea], m[eb], w0, w1)

    # --- watertight faces (gshell_tets.py:322-325) ---
    is1, is2 = ntri == 1, ntri == 2
    t1, t2 = int(is1.sum()), int(is2.sum())
    f1 = np.take_along_axis(idx6[is1], TRIANGLE_TABLE[code[is1]][:, :3], 1).reshape(-1, 3)
    f2 = np.take_along_axis(idx6[is2], TRIANGLE_TABLE[code[is2]][:, :6], 1).reshape(-1, 3)
    faces_wt = np.concatenate([f1, f2], 0)

    # --- normals / tangents on the watertight mesh (gshell_tets.py:9-78, 326-327) ---
    v_nrm = _auto_normals(verts, faces_wt)
    v_tng = _compute_tangents(verts, v_nrm, faces_wt, num_tets)

    # --- polygon loops: flat corner array [3*T1 | 4*T2] (gshell_tets.py:331-339) ---
    loop3 = np.take_along_axis(idx6[is1], MESH_EDGE_TABLE[code[is1]][:, :3], 1)  # (T1,3)
    loop4 = np.take_along_axis(idx6[is2], MESH_EDGE_TABLE[code[is2]][:, :4], 1)  # (T2,4)
    corners = np.concatenate([loop3.reshape(-1), loop4.reshape(-1)])
    nxt = np.concatenate([np.roll(loop3, -1, 1).reshape(-1), np.roll(loop4, -1, 1).reshape(-1)])
    npoly_corners = corners.shape[0]

    # --- boundary vertices, one per polygon edge i -> next(i) (gshell_tets.py:342-385) ---
    mi, mj = msdf_vert[corners], msdf_vert[nxt]
    nz = np.abs(np.sign(mi) + np.sign(mj)) != 2
    neg_mj = (-mj).astype(F32)
    big_d = (mi + neg_mj).astype(F32)
    nz &= np.abs(big_d) > EPS12
    safe_d = np.where(nz, big_d, F32(1))
    u0 = np.where(nz, (neg_mj / safe_d).astype(F32), F32(0)).astype(F32)
    u1 = np.where(nz, (mi / safe_d).astype(F32), F32(0)).astype(F32)
    bverts = _lerp2(verts[corners], verts[nxt], u0[:, None], u1[:, None])
    btng = _lerp2(v_tng[corners], v_tng[nxt], u0[:, None], u1[:, None])
    bmsdf = _lerp2(msdf_vert[corners], msdf_vert[nxt], u0, u1)
    verts_aug = np.concatenate([verts, bverts], 0)
    v_tng_aug = np.concatenate([v_tng, btng], 0)
    msdf_aug = np.concatenate([msdf_vert, bmsdf], 0)
    va = verts_aug.shape[0]

    # --- mSDF case index and cut triangulation in 6 buckets (gshell_tets.py:400-420) ---
    mo3 = (msdf_vert[loop3] > 0).astype(np.int64)
    mo4 = (msdf_vert[loop4] > 0).astype(np.int64)
    case3 = mo3 @ np.array([4, 2, 1]) if t1 else np.zeros(0, np.int64)
    case4 = mo4 @ np.array([8, 4, 2, 1]) if t2 else np.zeros(0, np.int64)
    loc3 = np.concatenate([loop3, nv + np.arange(3 * t1).reshape(-1, 3)], 1)
    loc4 = np.concatenate([loop4, nv + 3 * t1 + np.arange(4 * t2).reshape(-1, 4)], 1)
    n3, n4 = NUM_TRIANGLES_TRI_TABLE[case3], NUM_TRIANGLES_QUAD_TABLE[case4]
    buckets = []
    for k in (1, 2):
        sel = n3 == k
        buckets.append(np.take_along_axis(loc3[sel], TRIANGLE_TABLE_TRI[case3[sel]][:, :3 * k], 1).reshape(-1, 3))
    for k in (1, 2, 3, 4):
        sel = n4 == k
        buckets.append(np.take_along_axis(loc4[sel], TRIANGLE_TABLE_QUAD[case4[sel]][:, :3 * k], 1).reshape(-1, 3))
    faces_aug = np.concatenate(buckets, 0)
    bucket_counts = np.array([b.shape[0] for b in buckets], dtype=np.int64)

    # --- unreferenced rows of verts_aug are zeroed (gshell_tets.py:423-427) ---
    used = np.zeros(va, dtype=bool)
    used[faces_aug.reshape(-1)] = True
    verts_aug = verts_aug.copy()
    verts_aug[~used] = 0

    out = dict(
        verts_aug=verts_aug, faces_aug=faces_aug, v_tng_aug=v_tng_aug,
        n_verts_watertight=nv, vertices_watertight=verts, faces_watertight=faces_wt,
        v_tng_watertight=v_tng, msdf=msdf_aug, msdf_watertight=msdf_vert, msdf_boundary=msdf_aug[nv:],
        # integer intermediates for parity tests
        valid_ids=valid_ids, code=code, unique_edge_keys=ukey, n_unique_edges=int(ukey.shape[0]),
        edge_a=ea, edge_b=eb, corners=corners, t1=t1, t2=t2, fv=fv, bucket_counts=bucket_counts,
        used=used, v_nrm=v_nrm,
        # float intermediates for the adjoint
        _w0=w0, _w1=w1, _dd=dd, _u0=u0, _u1=u1, _nz=nz, _D=big_d, _nxt=nxt, _m=m, _pos=pos, _s=s,
        _msdf_sign=msdf_sign, _n_grid=n_grid, _num_tets=num_tets,
    )
    if not output_watertight_template:  # gshell_tets.py:440-445: only the three msdf keys survive in `extra`
        out["extra_keys"] = ("msdf", "msdf_watertight", "msdf_boundary")
    else:
        out["extra_keys"] = ("n_verts_watertight", "vertices_watertight", "faces_watertight", "v_tng_watertight",
                             "msdf", "msdf_watertight", "msdf_boundary")
    return out


def _auto_normals(verts, faces):
    """gshell_tets.py:9-34.  Sequential scatter order (all i0, then all i1, then all i2) like the CPU reference."""
    nv = verts.shape[0]
    v_nrm = np.zeros((nv, 3), F32)
    if faces.shape[0] == 0:
        return _finish_normals(v_nrm)
    v0, v1, v2 = verts[faces[:, 0]], verts[faces[:, 1]], verts[faces[:, 2]]
    a, b = (v1 - v0).astype(F32), (v2 - v0).astype(F32)
    if faces.shape[0] == 3:
        # quirk gshell_tets.py:19: torch.cross without dim takes the FIRST axis of size 3
        fn = _cross_f32(a, b, 0)  # fn[r, c]: crossed along the face axis r
    else:
        fn = _cross_f32(a, b, -1)
    for c in range(3):
        np.add.at(v_nrm, faces[:, c], fn)
    return _finish_normals(v_nrm)


def _finish_normals(v_nrm):
    d = np.sum((v_nrm * v_nrm).astype(F32), -1, keepdims=True, dtype=F32)
    v_nrm = np.where(d > F32(1e-20), v_nrm, np.array([0, 0, 1], F32))
    return _safe_normalize(v_nrm.astype(F32))


def _compute_tangents(verts, v_nrm, faces, num_tets):
    """gshell_tets.py:40-78 with v_tex indexed by vertex id (see vertex_uv)."""
    nv = verts.shape[0]
    tang_sum = np.zeros((nv, 3), F32)
    cnt = np.zeros((nv, 3), F32)
    if faces.shape[0]:
        p = [verts[faces[:, i]] for i in range(3)]
        t = [vertex_uv(faces[:, i], num_tets) for i in range(3)]
        uve1, uve2 = (t[1] - t[0]).astype(F32), (t[2] - t[0]).astype(F32)
        pe1, pe2 = (p[1] - p[0]).astype(F32), (p[2] - p[0]).astype(F32)
        nom = ((pe1 * uve2[:, 1:2]).astype(F32) - (pe2 * uve1[:, 1:2]).astype(F32)).astype(F32)
        den = ((uve1[:, 0:1] * uve2[:, 1:2]).astype(F32) - (uve1[:, 1:2] * uve2[:, 0:1]).astype(F32)).astype(F32)
        den = np.where(den > 0, np.maximum(den, F32(1e-6)), np.minimum(den, F32(-1e-6))).astype(F32)
        tang = (nom / den).astype(F32)
        for i in range(3):
            np.add.at(tang_sum, faces[:, i], tang)
            np.add.at(cnt, faces[:, i], F32(1))
    with np.errstate(invalid="ignore", divide="ignore"):
        tng = (tang_sum / cnt).astype(F32)
    tng = _safe_normalize(tng)
    proj = np.sum((tng * v_nrm).astype(F32), -1, keepdims=True, dtype=F32)
    return _safe_normalize((tng - (proj * v_nrm).astype(F32)).astype(F32))


def tangent_condition(fwd):
    """Per-row condition number of v_tng_aug (Va,) for the parity tests: how strongly a unit round-off in the ORDER of
    the scatter additions (sequential on the CPU reference, float atomics in arbitrary order on a GPU -- the
    reference's own CUDA scatter_add_ included) moves the unit tangent of a row.

    A watertight vertex's tangent is normalize(t - (t.n) n) with n = normalize(sum of face normals) and
    t = normalize(sum of per-face tangents / count) (gshell_tets.py:9-34, 40-78).  Re-ordering a sum S = sum x_f moves
    its direction by ~eps * sum|x_f| / |S|; the Gram-Schmidt step divides by s = |t - (t.n) n|.  So
        kappa = (sum|t_f| / |sum t_f| + 2 sum|n_f| / |sum n_f| + 3) / s,
    and a boundary row (lerp of its two corners' tangents with weights u0, u1, :380-385) inherits |u0| k_i + |u1| k_j.
    inf where a sum cancels exactly (the Kuhn lattice mixes left- and right-handed tets, so the reference's triangle
    table orients neighbouring faces inconsistently: such vertices are common on the synthetic grids)."""
    f8 = np.float64
    verts = fwd["vertices_watertight"].astype(f8)
    faces = fwd["faces_watertight"]
    nv = verts.shape[0]
    kap = np.full(nv, np.inf)
    if faces.shape[0] and faces.shape[0] != 3:      # (the 3-face torch.cross quirk is covered by a golden vector)
        p = [verts[faces[:, i]] for i in range(3)]
        t = [vertex_uv(faces[:, i], fwd["_num_tets"]).astype(f8) for i in range(3)]
        pe1, pe2 = p[1] - p[0], p[2] - p[0]
        uve1, uve2 = t[1] - t[0], t[2] - t[0]
        fn = np.cross(pe1, pe2)
        den = uve1[:, 0] * uve2[:, 1] - uve1[:, 1] * uve2[:, 0]
        den = np.where(den > 0, np.maximum(den, 1e-6), np.minimum(den, -1e-6))
        tang = (pe1 * uve2[:, 1:2] - pe2 * uve1[:, 1:2]) / den[:, None]
        sn, st = np.zeros((nv, 3)), np.zeros((nv, 3))
        an, at = np.zeros(nv), np.zeros(nv)
        for i in range(3):
            np.add.at(sn, faces[:, i], fn)
            np.add.at(st, faces[:, i], tang)
            np.add.at(an, faces[:, i], np.linalg.norm(fn, axis=1))
            np.add.at(at, faces[:, i], np.linalg.norm(tang, axis=1))
        ln, lt = np.linalg.norm(sn, axis=1), np.linalg.norm(st, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            n = sn / ln[:, None]
            tt = st / lt[:, None]
            g = tt - (tt * n).sum(-1, keepdims=True) * n
            s = np.linalg.norm(g, axis=1)
            kap = (at / lt + 2.0 * an / ln + 3.0) / s
        kap[~np.isfinite(kap)] = np.inf
        # the reference's thresholds (1e-20 on squared lengths) switch branches: treat anything near them as unstable
        kap[(ln * ln < 1e-18) | (lt * lt < 1e-18) | (s * s < 1e-18)] = np.inf
    corners, nxt = fwd["corners"], fwd["_nxt"]
    u0, u1 = np.abs(fwd["_u0"].astype(f8)), np.abs(fwd["_u1"].astype(f8))
    with np.errstate(invalid="ignore"):
        kb = np.where(u0 > 0, u0 * kap[corners], 0.0) + np.where(u1 > 0, u1 * kap[nxt], 0.0)
    return np.concatenate([kap, kb])


# --------------------------------------------------------------------------------------
# backward (hand-derived adjoint of the float pipeline)
# --------------------------------------------------------------------------------------
def tangent_backward(fwd, g_v_tng_aug=None, g_v_tng_watertight=None):
    """Adjoint of the tangent branch (gshell_tets.py:9-34 auto_normals, :40-78 compute_tangents, :380-385 the boundary
    interpolation of the tangents; SURVEY A.5 "optional branch"), float64.
    -> (g_vert (V,3): gradient w.r.t. the watertight vertex positions, g_mv (V,): w.r.t. msdf_vert through the boundary
    coefficients); both are ADDED to what extract_backward accumulates before the crossing-edge step.
    Not covered: a watertight mesh of exactly three faces (the torch.cross quirk, :19)."""
    f8 = np.float64
    nv = fwd["n_verts_watertight"]
    faces = fwd["faces_watertight"]
    if faces.shape[0] == 3:
        raise NotImplementedError("tangent gradients on a three-face mesh (torch.cross without dim)")
    verts = fwd["vertices_watertight"].astype(f8)
    corners, nxt = fwd["corners"], fwd["_nxt"]
    tng = fwd["v_tng_watertight"].astype(f8)
    g_t2 = np.zeros((nv, 3), f8)
    g_mv = np.zeros(nv, f8)
    if g_v_tng_watertight is not None:
        g_t2 += np.asarray(g_v_tng_watertight, f8)
    if g_v_tng_aug is not None:
        g_aug = np.asarray(g_v_tng_aug, f8)
        g_t2 += g_aug[:nv]
        g_bt = g_aug[nv:]                                  # (v_tng_aug rows are NOT zeroed for unused vertices, A.2.11)
        u0, u1, nz, big_d = fwd["_u0"].astype(f8), fwd["_u1"].astype(f8), fwd["_nz"], fwd["_D"].astype(f8)
        np.add.at(g_t2, corners, g_bt * u0[:, None])
        np.add.at(g_t2, nxt, g_bt * u1[:, None])
        g_u0 = np.sum(g_bt * tng[corners], -1)
        g_u1 = np.sum(g_bt * tng[nxt], -1)
        safe_d = np.where(nz, big_d, 1.0)
        g_d = np.where(nz, -(g_u0 * u0 + g_u1 * u1) / safe_d, 0.0)
        np.add.at(g_mv, corners, np.where(nz, g_u1 / safe_d + g_d, 0.0))
        np.add.at(g_mv, nxt, np.where(nz, -(g_u0 / safe_d + g_d), 0.0))
    g_vert = np.zeros((nv, 3), f8)
    if faces.shape[0] == 0:
        return g_vert, g_mv
    # ---- forward intermediates again, in float64 ----
    i0, i1, i2 = faces[:, 0], faces[:, 1], faces[:, 2]
    e1, e2 = verts[i1] - verts[i0], verts[i2] - verts[i0]
    fn = np.cross(e1, e2)
    n_sum = np.zeros((nv, 3), f8)
    for c in range(3):
        np.add.at(n_sum, faces[:, c], fn)
    uv = [vertex_uv(faces[:, i], fwd["_num_tets"]).astype(f8) for i in range(3)]
    uve1, uve2 = uv[1] - uv[0], uv[2] - uv[0]
    den = uve1[:, 0] * uve2[:, 1] - uve1[:, 1] * uve2[:, 0]
    den = np.where(den > 0, np.maximum(den, 1e-6), np.minimum(den, -1e-6))
    tang = (e1 * uve2[:, 1:2] - e2 * uve1[:, 1:2]) / den[:, None]
    t_sum = np.zeros((nv, 3), f8)
    cnt = np.zeros(nv, f8)
    for c in range(3):
        np.add.at(t_sum, faces[:, c], tang)
        np.add.at(cnt, faces[:, c], 1.0)

    def normalize_bwd(x, gy, eps=1e-20):
        """y = x / sqrt(max(x.x, eps)) -> (y, g_x)"""
        d = np.sum(x * x, -1, keepdims=True)
        length = np.sqrt(np.maximum(d, eps))
        y = x / length
        free = d > eps                                      # (clamped branch: the length is a constant)
        gx = np.where(free, (gy - y * np.sum(y * gy, -1, keepdims=True)) / length, gy / length)
        return y, gx

    nondeg = np.sum(n_sum * n_sum, -1, keepdims=True) > 1e-20
    n_in = np.where(nondeg, n_sum, np.array([0.0, 0.0, 1.0]))
    with np.errstate(invalid="ignore", divide="ignore"):
        a = t_sum / cnt[:, None]
    live = cnt > 0                                          # (a vertex of no face has a NaN tangent in the reference too)
    a = np.where(live[:, None], a, 0.0)
    n_unit, _ = normalize_bwd(n_in, np.zeros_like(n_in))
    t1, _ = normalize_bwd(a, np.zeros_like(a))
    proj = np.sum(t1 * n_unit, -1, keepdims=True)
    w = t1 - proj * n_unit
    _, g_w = normalize_bwd(w, g_t2)
    g_t1 = g_w - n_unit * np.sum(n_unit * g_w, -1, keepdims=True)
    g_n = -(proj * g_w + np.sum(n_unit * g_w, -1, keepdims=True) * t1)
    _, g_a = normalize_bwd(a, g_t1)
    g_s = np.where(live[:, None], g_a / np.where(live, cnt, 1.0)[:, None], 0.0)
    _, g_nin = normalize_bwd(n_in, g_n)
    g_nsum = np.where(nondeg, g_nin, 0.0)
    # ---- faces ----
    g_tang = g_s[i0] + g_s[i1] + g_s[i2]
    g_nom = g_tang / den[:, None]
    g_fn = g_nsum[i0] + g_nsum[i1] + g_nsum[i2]
    g_e1 = g_nom * uve2[:, 1:2] + np.cross(e2, g_fn)
    g_e2 = -g_nom * uve1[:, 1:2] + np.cross(g_fn, e1)
    np.add.at(g_vert, i1, g_e1)
    np.add.at(g_vert, i2, g_e2)
    np.add.at(g_vert, i0, -(g_e1 + g_e2))
    return g_vert, g_mv


def extract_backward(fwd, g_verts_aug=None, g_msdf=None, g_vertices_watertight=None, g_msdf_watertight=None,
                     g_v_tng_aug=None, g_v_tng_watertight=None, g_mvert_extra=None):
    """Gradients w.r.t. (pos, sdf, msdf) for upstream gradients on verts_aug, extra['msdf'],
    extra['vertices_watertight'], extra['msdf_watertight'].  float64 accumulation.

    Mirrors what autograd does through gshell_tets.py:291-303 (crossing interpolation, the stop-grad
    copy at :303), :342-397 (boundary interpolation; coefficients detached for the msdf attribute, :388-389)
    and the in-place zeroing at :427 (zeroed rows receive no gradient).
    g_v_tng_aug / g_v_tng_watertight: upstream gradients of the tangents (tangent_backward above).
    g_mvert_extra (V,): a gradient w.r.t. msdf_vert handed in directly (what d3h_backward_args.g_mvert_tng carries).
    """
    f8 = np.float64
    nv, n_grid = fwd["n_verts_watertight"], fwd["_n_grid"]
    corners, nxt = fwd["corners"], fwd["_nxt"]
    ncorn = corners.shape[0]
    va = nv + ncorn
    used = fwd["used"]
    verts = fwd["vertices_watertight"].astype(f8)
    g_va = np.zeros((va, 3), f8) if g_verts_aug is None else np.asarray(g_verts_aug, f8).copy()
    g_va[~used] = 0
    g_ma = np.zeros(va, f8) if g_msdf is None else np.asarray(g_msdf, f8)
    g_vert = g_va[:nv].copy()
    if g_vertices_watertight is not None:
        g_vert += np.asarray(g_vertices_watertight, f8)
    g_sg = g_ma[:nv].copy()  # grad wrt msdf_vert_stopvgd (weights detached)
    if g_msdf_watertight is not None:
        g_sg += np.asarray(g_msdf_watertight, f8)
    g_mv = np.zeros(nv, f8)  # grad wrt msdf_vert (through boundary coefficients)
    if g_mvert_extra is not None:
        g_mv += np.asarray(g_mvert_extra, f8)
    if g_v_tng_aug is not None or g_v_tng_watertight is not None:
        gt_vert, gt_mv = tangent_backward(fwd, g_v_tng_aug, g_v_tng_watertight)
        g_vert += gt_vert
        g_mv += gt_mv

    # boundary vertices
    u0, u1, nz, big_d = fwd["_u0"].astype(f8), fwd["_u1"].astype(f8), fwd["_nz"], fwd["_D"].astype(f8)
    g_b = g_va[nv:]
    g_bm = g_ma[nv:]
    np.add.at(g_vert, corners, g_b * u0[:, None])
    np.add.at(g_vert, nxt, g_b * u1[:, None])
    np.add.at(g_sg, corners, g_bm * u0)
    np.add.at(g_sg, nxt, g_bm * u1)
    g_u0 = np.sum(g_b * verts[corners], -1)
    g_u1 = np.sum(g_b * verts[nxt], -1)
    safe_d = np.where(nz, big_d, 1.0)
    g_d = np.where(nz, -(g_u0 * u0 + g_u1 * u1) / safe_d, 0.0)
    np.add.at(g_mv, corners, np.where(nz, g_u1 / safe_d + g_d, 0.0))
    np.add.at(g_mv, nxt, np.where(nz, -(g_u0 / safe_d + g_d), 0.0))

    # crossing edges
    ea, eb = fwd["edge_a"], fwd["edge_b"]
    w0, w1, dd = fwd["_w0"].astype(f8), fwd["_w1"].astype(f8), fwd["_dd"].astype(f8)
    pos, m = fwd["_pos"].astype(f8), fwd["_m"].astype(f8)
    g_pos = np.zeros((n_grid, 3), f8)
    g_sdf = np.zeros(n_grid, f8)
    g_m = np.zeros(n_grid, f8)
    np.add.at(g_pos, ea, g_vert * w0[:, None])
    np.add.at(g_pos, eb, g_vert * w1[:, None])
    g_w0 = np.sum(g_vert * pos[ea], -1) + g_mv * m[ea]
    g_w1 = np.sum(g_vert * pos[eb], -1) + g_mv * m[eb]
    np.add.at(g_m, ea, (g_mv + g_sg) * w0)
    np.add.at(g_m, eb, (g_mv + g_sg) * w1)
    g_dd = -(g_w0 * w0 + g_w1 * w1) / dd
    np.add.at(g_sdf, ea, g_w1 / dd + g_dd)
    np.add.at(g_sdf, eb, -(g_w0 / dd + g_dd))
    if fwd["_msdf_sign"] < 0:
        # hmsdf_tets_split.py:261-264 negates msdf under torch.no_grad(): type="body" never back-propagates into msdf
        g_m = None
    return g_pos, g_sdf, g_m
