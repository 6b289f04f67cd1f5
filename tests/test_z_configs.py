"""BASELINE.json configurations at their named sizes (size-independent properties + the surface counts the reference
itself produces), and the pipelined use of the batch API that bench.py relies on.
Runs last (file name) so that a failure here cannot hide the small-size parity tests."""
import numpy as np
import pytest
import torch

from d3human_code_b200 import grids
from tests import _util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


#: tests that also run on the edge-scan path with the other forms of the static edge list: "scan" lets the host choose
#: (run-length compressed on lattices, transposed rows on unstructured numberings), "scan_rows" forbids the compression,
#: "scan_csr" also the rows (edge_scan_kernel walking edge_b by CSR offsets)
_CSR_WALK_TESTS = ("test_cuda_matches_golden", "test_cuda_matches_oracle", "test_smplx_layout_split_extraction",
                   "test_random_tet_soups", "test_tet_soups_with_repeated_vertices")


@pytest.fixture(autouse=True, params=["sort", "static", "scan", "scan_rows", "scan_csr"])
def edges_mode(request):
    """Every test runs three times: on the general path (per-call radix sort + run-length scan of the crossing-edge keys),
    on the static edge table path (tet stream + bitmap over the grid's sorted edge list, built once per tet array) and on
    the edge-scan path (walk over the static edge list instead of the tet stream; the default of a training run).  A few
    run a fourth time on the edge-scan path with the CSR walk instead of the transposed edge rows."""
    from d3human_code_b200 import extract as E
    if request.param in ("scan_rows", "scan_csr") and request.node.originalname not in _CSR_WALK_TESTS:
        pytest.skip("the other forms of the static edge list are covered by the golden / oracle / soup tests")
    E.set_static_edges("0" if request.param == "sort" else "1")
    E.set_edge_scan(request.param.startswith("scan"))
    E.set_scan_rows(request.param != "scan_csr")
    E.set_scan_runs(request.param == "scan")
    yield "scan" if request.param.startswith("scan") else request.param
    E.set_static_edges("auto")
    E.set_edge_scan(True)
    E.set_scan_rows(True)
    E.set_scan_runs(True)


# counts measured by running the reference on CPU: Fv, V, Fw, Va, Fa
REFERENCE_COUNTS = {
    (64, "sphere"): (31608, 20702, 41400, 125318, 24708),
    (128, "sphere"): (127080, 83222, 166440, 503822, 97830),
    (64, "capsule"): (14162, 9392, 18780, 56496, 10412),
}


@pytest.mark.parametrize("res,field", sorted(REFERENCE_COUNTS))
def test_named_configs_reproduce_reference_counts(dev, res, field):
    """configs[0] (64^3 sphere + plane, the reference's own CPU-runnable case), its 128^3 version and the 64^3 capsule
    field: every count the reference reports, a closed manifold watertight mesh, and referenced rows == non-zero rows."""
    from d3human_code_b200.geometry.gshell_tets import GShell_Tets
    from d3human_code_b200.extract import last_counts
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = (grids.sphere_plane_field if field == "sphere" else grids.capsule_garment_field)(pos)
    verts, faces, _, _, v_tng, extra = GShell_Tets()(torch.tensor(pos, device=dev), torch.tensor(sdf, device=dev),
                                                     torch.tensor(msdf, device=dev), torch.tensor(tets, device=dev))
    c = last_counts()
    fv, v, fw, va, fa = REFERENCE_COUNTS[(res, field)]
    assert (c["n_valid_tets"], c["n_verts"], c["n_faces_watertight"], c["n_verts_aug"], c["n_faces_aug"]) == (fv, v, fw, va, fa)
    assert verts.shape == (va, 3) and faces.shape == (fa, 3) and extra["n_verts_watertight"] == v
    wt = extra["faces_watertight"]
    e = torch.sort(torch.cat([wt[:, [0, 1]], wt[:, [1, 2]], wt[:, [2, 0]]], 0), dim=1).values
    _, cnt = torch.unique(e[:, 0] * (v + 1) + e[:, 1], return_counts=True)
    assert bool((cnt == 2).all())                                   # closed 2-manifold
    used = torch.zeros(va, dtype=torch.bool, device=dev)
    used[faces.reshape(-1)] = True
    assert bool((verts[~used] == 0).all()) and bool((verts[used].abs().sum(1) > 0).all())
    assert int(faces.min()) >= 0 and int(faces.max()) < va
    # the open mesh keeps the part of the surface with positive mSDF: every kept watertight vertex has msdf > 0
    assert bool((extra["msdf_watertight"][used[:v]] > 0).all())


def test_pipelined_async_groups_equal_single_calls(dev):
    """bench.py's calling pattern: several extract_frames_async batches in flight at once (they queue up per lane, share
    the lane workspaces and the count ring), results read group by group, backward per group.  Every frame must equal
    the drop-in single call bit for bit; shared gradients must equal the sum over all frames."""
    from d3human_code_b200.extract import extract_frames_async
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    res, B, G = 24, 12, 3
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = grids.capsule_garment_field(pos)
    pos_b = np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in range(B)]).astype(np.float32)
    tt = torch.tensor(tets, device=dev)
    rng = np.random.default_rng(7)
    # ---- reference values: one drop-in call per frame ----
    ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    hm = hmSDF_Tets()
    singles, ups, gpos = [], [], []
    for f in range(B):
        tp = torch.tensor(pos_b[f], device=dev, requires_grad=True)
        verts, faces, _, _, _, extra = hm(tp, ts, tm, tt, "cloth")
        gv = torch.tensor(rng.standard_normal(tuple(verts.shape)).astype(np.float32), device=dev)
        gm = torch.tensor(rng.standard_normal(tuple(extra["msdf"].shape)).astype(np.float32), device=dev)
        torch.autograd.backward([verts, extra["msdf"]], [gv, gm])
        singles.append((verts.detach().clone(), faces.clone(), extra["msdf"].detach().clone(),
                        extra["faces_watertight"].clone()))
        ups.append((gv, gm))
        gpos.append(tp.grad.clone())
    want_sdf, want_msdf = ts.grad.clone(), tm.grad.clone()
    # ---- the same frames as G async groups, all launched before the first result is read; twice (steady state) ----
    for rep in range(2):
        ts2 = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
        tm2 = torch.tensor(msdf, device=dev, requires_grad=True)
        groups = [torch.tensor(pos_b[g * B // G:(g + 1) * B // G], device=dev, requires_grad=True) for g in range(G)]
        futs = [extract_frames_async(pg, ts2, tm2, tt, types="cloth", lanes=3) for pg in groups]
        for g, fut in enumerate(futs):
            outs = fut.result()
            lo = g * B // G
            for k, (verts, faces, _, _, _, extra) in enumerate(outs):
                sv, sf, sm, sw = singles[lo + k]
                assert torch.equal(verts.detach(), sv) and torch.equal(faces, sf), (rep, g, k)
                assert torch.equal(extra["msdf"].detach(), sm) and torch.equal(extra["faces_watertight"], sw)
            torch.autograd.backward([o[0] for o in outs] + [o[5]["msdf"] for o in outs],
                                    [ups[lo + k][0] for k in range(len(outs))] + [ups[lo + k][1] for k in range(len(outs))])
        for g, pg in enumerate(groups):
            lo = g * B // G
            for k in range(pg.shape[0]):
                U.assert_close_normwise(f"grad_pos[{lo + k}]", pg.grad[k].cpu().numpy(), gpos[lo + k].cpu().numpy(), U.GRAD_RTOL)
        U.assert_close_normwise("grad_sdf", ts2.grad.cpu().numpy(), want_sdf.cpu().numpy(), 2 * U.GRAD_RTOL)
        U.assert_close_normwise("grad_msdf", tm2.grad.cpu().numpy(), want_msdf.cpu().numpy(), 2 * U.GRAD_RTOL)


def test_split_pair_matches_two_calls(dev):
    """hmSDF_Tets.split (cloth + body of one iteration as one batch, train.py:1040-1047) == the two reference-style calls,
    bit for bit forward; gradients of the shared pos / sdf are the sums, msdf only gets the cloth part."""
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    g = grids.smplx_layout_grid(24, dilate=0.15, seed=2)       # configs[2] layout: scrambled, unstructured
    pos, tets = g["v"], g["f"]
    sdf, msdf = grids.capsule_garment_field(pos)
    hm = hmSDF_Tets()
    tt = torch.tensor(tets, device=dev)

    def leaves():
        return (torch.tensor(pos, device=dev, requires_grad=True), torch.tensor(sdf[:, None], device=dev, requires_grad=True),
                torch.tensor(msdf, device=dev, requires_grad=True))

    tp, ts, tm = leaves()
    cloth, body = hm.split(tp, ts, tm, tt)
    (cloth[0].square().sum() + body[0].square().sum() + cloth[5]["msdf"].sum() - body[5]["msdf"].sum()).backward()
    tp2, ts2, tm2 = leaves()
    c2 = hm(tp2, ts2, tm2, tt, "cloth")
    b2 = hm(tp2, ts2, tm2, tt, "body")
    (c2[0].square().sum() + b2[0].square().sum() + c2[5]["msdf"].sum() - b2[5]["msdf"].sum()).backward()
    for a, b in ((cloth, c2), (body, b2)):
        assert torch.equal(a[0].detach(), b[0].detach()) and torch.equal(a[1], b[1])
        assert torch.equal(a[5]["msdf"].detach(), b[5]["msdf"].detach())
        assert torch.equal(a[5]["faces_watertight"], b[5]["faces_watertight"])
    U.assert_close_normwise("grad_pos", tp.grad.cpu().numpy(), tp2.grad.cpu().numpy(), U.GRAD_RTOL)
    U.assert_close_normwise("grad_sdf", ts.grad.cpu().numpy(), ts2.grad.cpu().numpy(), U.GRAD_RTOL)
    U.assert_close_normwise("grad_msdf", tm.grad.cpu().numpy(), tm2.grad.cpu().numpy(), U.GRAD_RTOL)


@pytest.mark.parametrize("seed,n,f", [(0, 9, 60), (2, 25, 3000), (3, 12, 6000), (4, 300, 5000)])
def test_random_tet_soups(dev, seed, n, f):
    """Non-manifold tet soups (random vertex quadruples): edges shared by hundreds of tets, vertices with hundreds of
    neighbours -- the oversized-group path of the sort, long neighbour lists of the static edge table, dense buckets.
    tests/test_oracle_vs_reference.py shows the oracle tracks the reference on such input."""
    from oracle import gshell_oracle as O
    from tests import test_cuda_parity as G
    rng = np.random.default_rng(500 + seed)
    tets = np.stack([rng.permutation(n)[:4] for _ in range(f)]).astype(np.int64)
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    sdf = rng.standard_normal(n).astype(np.float32)
    msdf = rng.standard_normal(n).astype(np.float32)
    sdf[::7] = 0.0
    msdf[::5] = 0.0
    typ = (None, "cloth", "body")[seed % 3]
    cls = "GShell_Tets" if typ is None else "hmSDF_Tets"
    fwd = O.extract_forward(pos, sdf, msdf, tets, -1 if typ == "body" else 1, True)
    rng2 = np.random.default_rng(1)
    grads = dict(g_verts_aug=rng2.standard_normal(fwd["verts_aug"].shape).astype(np.float32),
                 g_msdf=rng2.standard_normal(fwd["msdf"].shape).astype(np.float32),
                 g_msdf_watertight=None, g_vertices_watertight=None)
    out, g = G._run(dev, pos, sdf, msdf, tets, cls, typ, True, grads)
    U.assert_exact("faces_aug", out["faces_aug"], fwd["faces_aug"])
    U.assert_exact("verts_aug", out["verts_aug"], fwd["verts_aug"])
    U.assert_exact("msdf", out["msdf"], fwd["msdf"])
    U.assert_exact("faces_watertight", out["faces_watertight"], fwd["faces_watertight"])
    g_pos, g_sdf, g_msdf = O.extract_backward(fwd, grads["g_verts_aug"], grads["g_msdf"])
    U.assert_close_normwise("grad_pos", g[0], g_pos, 5 * U.GRAD_RTOL)
    U.assert_close_normwise("grad_sdf", g[1], g_sdf, 5 * U.GRAD_RTOL)
    if typ != "body":
        U.assert_close_normwise("grad_msdf", g[2], g_msdf, 5 * U.GRAD_RTOL)


@pytest.mark.parametrize("seed,n,f,wt", [(0, 10, 400, True), (1, 30, 2500, True), (2, 16, 1500, False)])
def test_tet_soups_with_repeated_vertices(dev, seed, n, f, wt):
    """Tets drawn WITH replacement: some list a vertex twice (or more), so a tet meets one of its edges twice and has
    self-edges (v,v).  The marking kernel elects the thread of a tet's first crossing edge by rule instead of by atomic;
    this is the input where that rule has to agree with the reference's plain enumeration.  Both edge paths (fixture)."""
    from oracle import gshell_oracle as O
    from tests import test_cuda_parity as G
    rng = np.random.default_rng(900 + seed)
    tets = rng.integers(0, n, size=(f, 4)).astype(np.int64)
    assert (np.sort(tets, 1)[:, 1:] == np.sort(tets, 1)[:, :-1]).any()
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    sdf = rng.standard_normal(n).astype(np.float32)
    msdf = rng.standard_normal(n).astype(np.float32)
    sdf[::6] = 0.0
    fwd = O.extract_forward(pos, sdf, msdf, tets, 1, wt)
    rng2 = np.random.default_rng(1)
    grads = dict(g_verts_aug=rng2.standard_normal(fwd["verts_aug"].shape).astype(np.float32),
                 g_msdf=rng2.standard_normal(fwd["msdf"].shape).astype(np.float32),
                 g_msdf_watertight=None, g_vertices_watertight=None)
    out, g = G._run(dev, pos, sdf, msdf, tets, "GShell_Tets", None, wt, grads)
    U.assert_exact("faces_aug", out["faces_aug"], fwd["faces_aug"])
    U.assert_exact("verts_aug", out["verts_aug"], fwd["verts_aug"])
    U.assert_exact("msdf", out["msdf"], fwd["msdf"])
    if wt:
        U.assert_exact("faces_watertight", out["faces_watertight"], fwd["faces_watertight"])
    g_pos, g_sdf, g_msdf = O.extract_backward(fwd, grads["g_verts_aug"], grads["g_msdf"])
    U.assert_close_normwise("grad_pos", g[0], g_pos, 5 * U.GRAD_RTOL)
    U.assert_close_normwise("grad_sdf", g[1], g_sdf, 5 * U.GRAD_RTOL)
    U.assert_close_normwise("grad_msdf", g[2], g_msdf, 5 * U.GRAD_RTOL)


def fused_pair_equals_two_calls(dev, res, field):      # collected by tests/test_zzzz_fused_pair.py (sorts last: opt-in path)
    """hmSDF_Tets.split(fused=True): one classification / edge de-duplication / vertex interpolation for
    the cloth / body pair of an iteration, only the mSDF cut is replayed for the body.  Bit-identical to the two separate
    calls; gradients equal to their sum.  Developed on the CPU emulation of the kernels (tests/test_emu_parity.py); this is
    its check on the real GPU -- run as the very last test of the suite (tests/test_zzzz_fused_pair.py)."""
    from tests import test_cuda_parity as G
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    pos, sdf, msdf, tets = G._inputs(res, field, seed=res)
    hm = hmSDF_Tets()
    tt = torch.tensor(tets, device=dev)

    def leaves():
        return (torch.tensor(pos, device=dev, requires_grad=True), torch.tensor(sdf, device=dev, requires_grad=True),
                torch.tensor(msdf, device=dev, requires_grad=True))

    def loss(c, b):
        return (c[0].square().sum() + b[0].square().sum() + c[5]["msdf"].sum() - 2.0 * b[5]["msdf"].sum()
                + c[5]["vertices_watertight"].sum() + (b[5]["msdf_watertight"] ** 2).sum())

    for rep in range(2):
        tp, ts, tm = leaves()
        cloth, body = hm.split(tp, ts, tm, tt, fused=True)
        loss(cloth, body).backward()
        tp2, ts2, tm2 = leaves()
        c2 = hm(tp2, ts2, tm2, tt, "cloth")
        b2 = hm(tp2, ts2, tm2, tt, "body")
        loss(c2, b2).backward()
        for a, b in ((cloth, c2), (body, b2)):
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), rep
            for k in a[5]:
                if torch.is_tensor(a[5][k]) and "tng" not in k:
                    assert torch.equal(a[5][k], b[5][k]), (rep, k)
        U.assert_close_normwise("grad_pos", tp.grad.cpu().numpy(), tp2.grad.cpu().numpy(), U.GRAD_RTOL)
        U.assert_close_normwise("grad_sdf", ts.grad.cpu().numpy(), ts2.grad.cpu().numpy(), U.GRAD_RTOL)
        U.assert_close_normwise("grad_msdf", tm.grad.cpu().numpy(), tm2.grad.cpu().numpy(), U.GRAD_RTOL)
