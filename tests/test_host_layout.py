"""Host logic of the batch path without a GPU: the argument blocks of a batch are an int64 matrix filled with vectorised
numpy assignments (extract._Layout); here every row is re-read through the ctypes mirror of d3h_forward_args and the
slab geometry is checked for overlaps, alignment and lane / workspace assignment."""
import ctypes as C

import numpy as np
import pytest
import torch

from d3human_code_b200 import _cabi
from d3human_code_b200 import extract as E


def _plan(cap_tets=1000, cap_v=700, cap_va=4100, cap_fw=1500, cap_fa=900, lanes=3):
    p = E._Plan(device=torch.device("cpu"), n_tets=6 * 16 ** 3, n_grid=17 ** 3)
    p.cap_tets, p.cap_v, p.cap_va, p.cap_fw, p.cap_fa = cap_tets, cap_v, cap_va, cap_fw, cap_fa
    p.workspace_bytes = 1 << 20
    p.workspace_ptrs = [0x10000000 + 0x100000 * i for i in range(lanes)]
    p.counts_ptr = 0x7f0000000000
    return p


def _rows(lay, bases):
    A = lay.A.copy()
    np.add(lay.OFF, np.array([bases[k] for k in lay.slab_of], dtype=np.int64), out=A[:, lay.c0:lay.c1])
    if lay.static is not None:
        A[:, E._FC["vacc"]] = lay.vacc_off + bases[0]
    return [_cabi.ForwardArgs.from_buffer_copy(A[i].tobytes()) for i in range(A.shape[0])]


@pytest.mark.parametrize("static", [False, True])
def test_layout_rows_are_valid_forward_args(static):
    B, lanes = 5, 3
    plan = _plan(lanes=lanes)
    st = None
    if static:
        st = (torch.zeros(plan.n_grid + 1, dtype=torch.int32), torch.zeros((40000, 2), dtype=torch.int32), 40000)
    lay = E._Layout(plan, B, lanes, 1, st)
    bases = (0x20000000, 0x30000000, 0x40000000)   # float / int64 / int32 slab
    rows = _rows(lay, bases)
    cv, cva, cfw, cfa, ct = lay.caps
    assert (cv, cva, cfw, cfa, ct) == (700, 4100, 1500, 900, 1000)
    for i, a in enumerate(rows):
        assert (a.n_grid, a.n_tets, a.tet_begin, a.tet_end) == (plan.n_grid, plan.n_tets, 0, plan.n_tets)
        assert (a.cap_valid_tets, a.cap_verts, a.cap_verts_aug, a.cap_faces_wt, a.cap_faces_aug) == (ct, cv, cva, cfw, cfa)
        assert a.workspace == plan.workspace_ptrs[i % lanes] and a.workspace_bytes == plan.workspace_bytes
        f0, f1 = bases[0] + 4 * i * lay.f_len, bases[0] + 4 * (i + 1) * lay.f_len
        i0, i1 = bases[1] + 8 * i * lay.i_len, bases[1] + 8 * (i + 1) * lay.i_len
        t0, t1 = bases[2] + 4 * i * lay.t_len, bases[2] + 4 * (i + 1) * lay.t_len
        # every region lies inside the frame's slice of its slab, 16-byte aligned, and the regions do not overlap
        fregs = [(a.verts_aug, 12 * cva), (a.v_tng_aug, 12 * cva), (a.msdf_aug, 4 * cva), (a.verts_wt, 12 * cv),
                 (a.v_tng_wt, 12 * cv), (a.msdf_wt, 4 * cv)]
        if static:
            fregs.append((a.vacc, 32 * cv))
            assert a.edge_off == st[0].data_ptr() and a.edge_ab == st[1].data_ptr() and a.n_edges == 40000
        else:
            assert not a.edge_off and not a.vacc and a.n_edges == 0
        iregs = [(a.faces_aug, 24 * cfa), (a.faces_wt, 24 * cfw)]
        tregs = [(a.tape_edges, 8 * cv), (a.tape_corners, 16 * ct), (a.tape_slots, 16 * ct), (a.tape_runs, 4 * (cv + 1))]
        for regs, (lo, hi) in ((fregs, (f0, f1)), (iregs, (i0, i1)), (tregs, (t0, t1))):
            spans = sorted((p, p + n) for p, n in regs)
            assert spans[0][0] >= lo and spans[-1][1] <= hi, (i, spans, lo, hi)
            for (p0, e0), (p1, _) in zip(spans, spans[1:]):
                assert e0 <= p1, "regions overlap"
            for p, _ in regs:
                assert p % 16 == 0
    # frames on different lanes never share a workspace, frames of one lane do
    ws = [a.workspace for a in rows]
    assert ws[0] == ws[3] and ws[1] == ws[4] and len({ws[0], ws[1], ws[2]}) == 3


def test_column_maps_match_ctypes_structs():
    for cols, words, st in ((E._FC, E._FW, _cabi.ForwardArgs), (E._BC, E._BW, _cabi.BackwardArgs), (E._CC, E._CW, _cabi.Counts)):
        assert words * 8 == C.sizeof(st)
        for name, _ in st._fields_:
            assert cols[name] == getattr(st, name).offset // 8
    # the two int32 flags of each block share one word: low half first (little endian)
    a = _cabi.ForwardArgs()
    a.msdf_negate, a.watertight_template = 1, 1
    word = np.frombuffer(bytes(a), dtype=np.int64)[E._FC["msdf_negate"]]
    assert word == 1 | (1 << 32)
    b = _cabi.BackwardArgs()
    b.msdf_negate, b.grads_prezeroed = 0, 1
    assert np.frombuffer(bytes(b), dtype=np.int64)[E._BC["msdf_negate"]] == (1 << 32)


def test_capacity_prediction_is_stable():
    for need in (0, 1, 1000, 57740, 10 ** 6):
        g = E._grow(need)
        assert g >= need + 1024 and E._shrink(g, need) == g
        assert E._shrink(g, int(need * 1.05)) >= int(need * 1.05)       # a little growth: capacity follows
        assert E._shrink(10 * g + 5000, need) == g                       # a collapsed surface: capacity shrinks


def test_static_edge_policy_without_gpu():
    E.set_static_edges("0")
    try:
        assert E.static_edges_for(torch.zeros((4, 4), dtype=torch.int32), 8) is None
        with pytest.raises(ValueError):
            E.set_static_edges("maybe")
    finally:
        E.set_static_edges("auto")


def test_build_edge_table_on_cpu_matches_numpy():
    """The one-time edge table build is plain torch: it runs on CPU tensors too."""
    from d3human_code_b200 import grids
    pos, tets = grids.kuhn_grid(6)
    n = pos.shape[0]
    E.set_edge_scan(False)
    try:
        off, ab, u, tet_rank, edge_b, etet_off, etets, etets8, rows, row_off, eruns, truns = E.build_edge_table(torch.tensor(tets, dtype=torch.int32), n)
        assert tet_rank is None and edge_b is None and etets is None and rows is None and row_off is None and eruns[0] is None and truns[0] is None
        E.set_tet_edge_ranks(True)
        _, _, _, tet_rank, edge_b, *_ = E.build_edge_table(torch.tensor(tets, dtype=torch.int32), n)
        assert edge_b is None
    finally:
        E.set_tet_edge_ranks(False)
        E.set_edge_scan(True)
    # the tables of the edge-scan path: larger endpoints + the tets around every edge
    E.set_mark_rows(True)
    E.set_scan_rows(False)
    E.set_scan_runs(False)
    try:
        _, ab2, u2, tet_rank2, edge_b, etet_off, etets, etets8, rows, _, (runs, _, _), _ = E.build_edge_table(torch.tensor(tets, dtype=torch.int32), n)
    finally:
        E.set_scan_rows(True)
        E.set_scan_runs(True)
    assert rows is None and runs is None
    assert u2 == u and torch.equal(ab2, ab) and torch.equal(tet_rank2, tet_rank)
    assert torch.equal(edge_b, ab[:, 1]) and edge_b.is_contiguous() and edge_b.dtype == torch.int32
    eo, et = etet_off.numpy(), etets.numpy()
    assert eo[0] == 0 and eo[-1] == 6 * tets.shape[0] and etets.dtype == torch.int32
    want = [set() for _ in range(u)]
    for t, row in enumerate(tet_rank.numpy()[:, :6]):
        for r in row:
            want[r].add(t)
    for r in (0, 1, u // 2, u - 1):
        seg = et[eo[r]:eo[r + 1]]
        assert set(seg.tolist()) == want[r] and np.all(np.diff(seg) >= 0)
    assert all(eo[r + 1] - eo[r] == len(want[r]) for r in range(u))    # (no tet of a lattice repeats an edge)
    # fixed-width incidence rows: the first 8 tets of every edge, -1 padding (a Kuhn lattice has at most 8 tets per edge)
    e8 = etets8.numpy()
    assert e8.shape == (u, 8) and etets8.dtype == torch.int32 and not (e8 == -2).any()
    for r in (0, 1, u // 2, u - 1):
        k = eo[r + 1] - eo[r]
        assert np.array_equal(e8[r, :k], et[eo[r]:eo[r + 1]]) and (e8[r, k:] == -1).all()
    E.set_mark_rows(False)
    assert tet_rank.shape == (tets.shape[0], 8) and tet_rank.dtype == torch.int32
    pairs = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))
    for e, (i, j) in enumerate(pairs):     # the rank of every tet edge points back at its endpoints
        got = ab.numpy()[tet_rank.numpy()[:, e]]
        assert np.array_equal(got[:, 0], np.minimum(tets[:, i], tets[:, j])) and np.array_equal(got[:, 1], np.maximum(tets[:, i], tets[:, j]))
    ea = np.minimum(tets[:, [0, 0, 0, 1, 1, 2]], tets[:, [1, 2, 3, 2, 3, 3]]).reshape(-1).astype(np.int64)
    eb = np.maximum(tets[:, [0, 0, 0, 1, 1, 2]], tets[:, [1, 2, 3, 2, 3, 3]]).reshape(-1).astype(np.int64)
    uk = np.unique(ea * n + eb)
    assert u == uk.shape[0]
    assert np.array_equal(ab.numpy().astype(np.int64), np.stack([uk // n, uk % n], 1))
    o = off.numpy()
    assert o[0] == 0 and o[-1] == u
    for a in (0, 7, n - 2):
        assert np.array_equal(ab.numpy()[o[a]:o[a + 1], 0], np.full(o[a + 1] - o[a], a))


def test_edge_table_incidence_of_degenerate_and_crowded_edges():
    """A tet that repeats a vertex meets one of its edges twice (listed once); an edge with more than 8 tets is flagged
    in the fixed-width rows (-2 in the last slot) and keeps its full list in the CSR form."""
    n = 24
    fan = [[0, 1, 2 + k, 3 + k] for k in range(10)]            # edge (0,1) is shared by 10 tets
    tets = np.array(fan + [[14, 14, 15, 16], [17, 18, 17, 19]], dtype=np.int32)
    E.set_mark_rows(True)
    _, ab, u, tet_rank, edge_b, etet_off, etets, etets8, *_ = E.build_edge_table(torch.tensor(tets), n)
    ab, eo, et, e8 = ab.numpy(), etet_off.numpy(), etets.numpy(), etets8.numpy()
    r01 = int(np.flatnonzero((ab[:, 0] == 0) & (ab[:, 1] == 1))[0])
    assert np.array_equal(et[eo[r01]:eo[r01 + 1]], np.arange(10)) and e8[r01, 7] == -2 and np.array_equal(e8[r01, :7], np.arange(7))
    r45 = int(np.flatnonzero((ab[:, 0] == 14) & (ab[:, 1] == 15))[0])           # edges (0,2) and (1,2) of tet 10
    assert np.array_equal(et[eo[r45]:eo[r45 + 1]], [10]) and np.array_equal(e8[r45], [10, -1, -1, -1, -1, -1, -1, -1])
    r78 = int(np.flatnonzero((ab[:, 0] == 17) & (ab[:, 1] == 18))[0])           # edges (0,1) and (1,2) of tet 11
    assert np.array_equal(et[eo[r78]:eo[r78 + 1]], [11])
    assert eo[-1] == et.shape[0] < 6 * tets.shape[0]
    E.set_mark_rows(False)


def _check_edge_rows(off, ab, rows, row_off, n):
    """Every slot of the transposed rows against the CSR list it restates (include/d3h_tets.h: edge_rows)."""
    off, ab, rows, row_off = off.numpy().astype(np.int64), ab.numpy(), rows.numpy(), row_off.numpy().astype(np.int64)
    n_chunks = (n + 31) // 32
    assert row_off.shape == (n_chunks + 1,) and row_off[0] == 0 and rows.shape == (32 * (row_off[-1] + 8),)
    assert (rows[32 * row_off[-1]:] == 0).all()                   # the spare rows
    deg = off[1:] - off[:-1]
    for c in range(n_chunks):
        lo, hi = 32 * c, min(32 * c + 32, n)
        w = row_off[c + 1] - row_off[c]
        assert w == deg[lo:hi].max()
        if w == 0:
            continue
        blk = rows[32 * row_off[c]:32 * row_off[c + 1]].reshape(w, 32)
        for l in range(32):
            v = lo + l
            if v >= n:
                assert (blk[:, l] == 0).all()
                continue
            d = deg[v]
            assert np.array_equal(blk[:d, l], ab[off[v]:off[v] + d, 1]) and (blk[d:, l] == v).all()


def test_edge_rows_restates_the_csr_list():
    """The transposed edge rows of the stream kernel: lattice (7 larger neighbours inside, fewer at the faces, a grid size
    that is no multiple of 32) and a small irregular soup with a vertex of degree > 8 and isolated vertices."""
    from d3human_code_b200 import grids
    pos, tets = grids.kuhn_grid(6)
    n = pos.shape[0]
    assert n % 32 != 0
    E.set_scan_runs(False)
    try:
        off, ab, u, _, edge_b, _, _, _, rows, row_off, (runs, _, _), _ = E.build_edge_table(torch.tensor(tets, dtype=torch.int32), n)
    finally:
        E.set_scan_runs(True)
    assert runs is None
    assert edge_b is None and rows.dtype == torch.int32 and row_off.dtype == torch.int32 and rows.is_contiguous()
    _check_edge_rows(off, ab, rows, row_off, n)
    assert int(row_off[-1]) * 32 >= u
    n = 70
    fan = [[0, 1 + k, 2 + k, 3 + k] for k in range(0, 30, 2)]      # vertex 0: 32 larger neighbours... a long chunk
    tets = np.array(fan + [[40, 41, 42, 43], [64, 65, 66, 69]], dtype=np.int32)
    off, ab, u, _, _, _, _, _, rows, row_off, (runs, _, _), _ = E.build_edge_table(torch.tensor(tets), n)
    assert runs is None                      # an irregular soup: too few edges per entry, the rows are built instead
    _check_edge_rows(off, ab, rows, row_off, n)
    assert int(row_off[1] - row_off[0]) > 8


def _check_edge_runs(ab, runs, chunk, ids, n):
    """The run-length compressed edge list against the sorted list it restates (include/d3h_tets.h: edge_runs): entries
    ascend in (chunk, d), masks and id rows name exactly the grid's edges."""
    ab, runs, chunk, ids = ab.numpy().astype(np.int64), runs.numpy().astype(np.int64), chunk.numpy().astype(np.int64), ids.numpy()
    assert chunk.shape == (runs.shape[0],) and ids.shape == (runs.shape[0], 32)
    key = chunk * n + runs[:, 0]
    assert np.all(np.diff(key) > 0) and np.all(runs[:, 0] >= 0)
    seen = np.zeros(ab.shape[0], dtype=bool)
    for k in range(runs.shape[0]):
        d, m = runs[k]
        m &= 0xFFFFFFFF
        assert m != 0
        for l in range(32):
            e = ids[k, l]
            assert (e >= 0) == bool((m >> l) & 1)
            if e >= 0:
                v = 32 * chunk[k] + l
                assert v < n and ab[e, 0] == v and ab[e, 1] == v + d and not seen[e]
                seen[e] = True
    assert seen.all()


def test_edge_runs_restate_the_csr_list():
    """Lattice (a grid size that is no multiple of 32; lane 31 of a mask is the int32 sign bit) and a tet soup with self
    edges (tets that repeat a vertex), forced through the compression."""
    from d3human_code_b200 import grids
    pos, tets = grids.kuhn_grid(6)
    n = pos.shape[0]
    off, ab, u, _, edge_b, _, _, _, rows, _, (runs, run_chunk, run_ids), (truns, trun_chunk, trun_ids) = E.build_edge_table(torch.tensor(tets, dtype=torch.int32), n)
    assert edge_b is None and rows is None and runs.dtype == torch.int32 and runs.shape[1] == 2 and runs.is_contiguous()
    assert runs.shape[0] * 4 <= u
    assert (runs[:, 1] < 0).any()
    _check_edge_runs(ab, runs, run_chunk, run_ids, n)
    _check_tet_runs(tets, truns, trun_chunk, trun_ids, n)
    assert truns.shape[0] * 4 <= tets.shape[0]
    rng = np.random.default_rng(5)
    n = 100
    tets = rng.integers(0, n, size=(300, 4)).astype(np.int32)
    tets[::7, 1] = tets[::7, 0]                                   # self edges
    E.set_scan_runs(False)
    try:
        off, ab, u, *_ = E.build_edge_table(torch.tensor(tets), n)
    finally:
        E.set_scan_runs(True)
    assert (ab[:, 0] == ab[:, 1]).any()
    runs, run_chunk, run_ids = E.build_edge_runs(off, ab, n, min_edges_per_entry=0)
    _check_edge_runs(ab, runs, run_chunk, run_ids, n)
    # a grid without structure in its numbering: hardly any entry is shared by two lanes -> not worth it
    tets = rng.integers(0, 5000, size=(300, 4)).astype(np.int32)
    off, ab, *_ = E.build_edge_table(torch.tensor(tets), 5000)
    assert E.build_edge_runs(off, ab, 5000) == (None, None, None)


def _check_tet_runs(tets, runs, chunk, ids, n):
    """The compressed tet array against the tets it restates (include/d3h_tets.h: tet_runs)."""
    runs, chunk, ids = runs.numpy().astype(np.int64), chunk.numpy().astype(np.int64), ids.numpy()
    assert chunk.shape == (runs.shape[0],) and ids.shape == (runs.shape[0], 32) and np.all(np.diff(chunk) >= 0)
    seen = np.zeros(tets.shape[0], dtype=bool)
    for k in range(runs.shape[0]):
        d1, d2, d3, m = runs[k]
        m &= 0xFFFFFFFF
        assert m != 0
        for l in range(32):
            t = ids[k, l]
            assert (t >= 0) == bool((m >> l) & 1)
            if t >= 0:
                v0 = 32 * chunk[k] + l
                assert tuple(tets[t]) == (v0, v0 + d1, v0 + d2, v0 + d3) and not seen[t]
                seen[t] = True
    assert seen.all()


def test_tet_runs_of_soups_and_refusals():
    """Shapes with negative differences and repeated vertices; duplicates of a tet and numberings without structure
    are refused (the marking kernel walks the incidence lists then)."""
    rng = np.random.default_rng(11)
    n = 64
    tets = rng.integers(0, n, size=(4000, 4)).astype(np.int32)
    tets = np.unique(tets, axis=0)
    tets[::9, 2] = tets[::9, 0]
    tets = np.unique(tets, axis=0)
    runs, chunk, ids = E.build_tet_runs(torch.tensor(tets), n, min_tets_per_entry=0)
    assert (runs[:, :3] < 0).any()
    _check_tet_runs(tets, runs, chunk, ids, n)
    assert E.build_tet_runs(torch.tensor(np.concatenate([tets, tets[:1]])), n, min_tets_per_entry=0) == (None, None, None)
    assert E.build_tet_runs(torch.tensor(tets), n) == (None, None, None)
