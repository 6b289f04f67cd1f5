"""Parity of the CUDA path (through the drop-in classes -> autograd.Function -> C ABI -> sm_100a kernels) with
  (1) the committed golden vectors produced by the live reference, and
  (2) the numpy oracle on seeded inputs at sizes it finishes in seconds,
plus size-independent properties at the BASELINE.json full size (128^3).

Tolerances (BASELINE.json north_star): topology / indices / counts / ordering bit-exact, positions 1e-6 relative
(asserted bit-exact), gradients 1e-5 normwise, tangents: per-row bound 2e-6 + 4e-7 * kappa from the condition of the
row's scatter sums (float-atomic order; oracle.tangent_condition, tests/_util.py).
"""
import numpy as np
import pytest
import torch

from oracle import gshell_oracle as O
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


def _classes():
    from d3human_code_b200.geometry.gshell_tets import GShell_Tets
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    return GShell_Tets(), hmSDF_Tets()


def _run(dev, pos, sdf, msdf, tets, cls="GShell_Tets", typ=None, wt=True, grads=None):
    g, h = _classes()
    tp = torch.tensor(pos, device=dev, requires_grad=True)
    ts = torch.tensor(sdf, device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    tt = torch.tensor(tets, device=dev)
    if cls == "GShell_Tets":
        verts, faces, uvs, uv_idx, v_tng, extra = g(tp, ts, tm, tt, wt)
    else:
        verts, faces, uvs, uv_idx, v_tng, extra = h(tp, ts, tm, tt, typ, wt)
    assert uvs is None and uv_idx is None
    assert faces.dtype == torch.int64 and verts.dtype == torch.float32
    out = dict(verts_aug=verts.detach().cpu().numpy(), faces_aug=faces.cpu().numpy(),
               v_tng_aug=v_tng.detach().cpu().numpy())
    for k, v in extra.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else v
    out["extra_keys"] = tuple(extra.keys())
    res = None
    if grads is not None:
        loss = (verts * torch.tensor(grads["g_verts_aug"], device=dev)).sum() \
            + (extra["msdf"] * torch.tensor(grads["g_msdf"], device=dev)).sum()
        if grads.get("g_msdf_watertight") is not None:
            loss = loss + (extra["msdf_watertight"] * torch.tensor(grads["g_msdf_watertight"], device=dev)).sum()
        if grads.get("g_vertices_watertight") is not None:
            loss = loss + (extra["vertices_watertight"] * torch.tensor(grads["g_vertices_watertight"], device=dev)).sum()
        if loss.requires_grad:
            loss.backward()
        res = tuple(None if t.grad is None else t.grad.cpu().numpy() for t in (tp, ts, tm))
    return out, res


# ---------------------------------------------------------------------------------------------- golden
@pytest.mark.parametrize("name", U.golden_cases())
def test_cuda_matches_golden(dev, name):
    rec = U.load_golden(name)
    grads = {k: rec.get(k) for k in ("g_verts_aug", "g_msdf", "g_msdf_watertight", "g_vertices_watertight")}
    out, g = _run(dev, rec["pos"], rec["sdf"], rec["msdf"], rec["tets"], rec["cls"], rec["type"], rec["wt"], grads)
    # tangents: condition-aware bound from the oracle's forward on the golden inputs (the exactly-three-faces fixture
    # exercises the torch.cross quirk, where the bound does not apply: flat tolerance there)
    fwd = O.extract_forward(rec["pos"], rec["sdf"], rec["msdf"], rec["tets"], rec["sign"], rec["wt"])
    kappa = O.tangent_condition(fwd) if fwd["faces_watertight"].shape[0] != 3 else None
    U.check_forward_against_golden(out, rec, tng_atol=U.TNG_CUDA_ATOL, tng_p99=U.TNG_CUDA_P99, kappa=kappa)
    if "grad_pos" in rec:
        U.check_grads_against_golden(g[0], g[1], g[2], rec)
    else:
        assert out["verts_aug"].shape[0] == 0


# ---------------------------------------------------------------------------------------------- oracle
def _inputs(res, field, seed=0):
    pos, tets = grids.kuhn_grid(res)
    if field == "sphere":
        sdf, msdf = grids.sphere_plane_field(pos)
    elif field == "capsule":
        sdf, msdf = grids.capsule_garment_field(pos)
    else:
        pos, sdf, msdf = grids.adversarial_field(pos, res, seed)
    return pos, sdf, msdf, tets


@pytest.mark.parametrize("res,field,cls,typ,wt", [
    (32, "sphere", "GShell_Tets", None, True),
    (48, "capsule", "hmSDF_Tets", "cloth", True),
    (48, "capsule", "hmSDF_Tets", "body", True),
    (20, "adv", "GShell_Tets", None, True),
    (20, "adv", "hmSDF_Tets", "body", True),
    (16, "adv", "GShell_Tets", None, False),
    (16, "adv", "hmSDF_Tets", "body", False),
    (33, "adv", "hmSDF_Tets", "cloth", True),   # > 1M corners: many sort / rle / poly tiles, look-back chains
])
def test_cuda_matches_oracle(dev, res, field, cls, typ, wt):
    pos, sdf, msdf, tets = _inputs(res, field, seed=res)
    sign = -1 if typ == "body" else 1
    fwd = O.extract_forward(pos, sdf, msdf, tets, sign, wt)
    rng = np.random.default_rng(5)
    grads = dict(g_verts_aug=rng.standard_normal(fwd["verts_aug"].shape).astype(np.float32),
                 g_msdf=rng.standard_normal(fwd["msdf"].shape).astype(np.float32),
                 g_msdf_watertight=rng.standard_normal(fwd["msdf_watertight"].shape).astype(np.float32),
                 g_vertices_watertight=(rng.standard_normal(fwd["vertices_watertight"].shape).astype(np.float32) if wt else None))
    out, g = _run(dev, pos, sdf, msdf, tets, cls, typ, wt, grads)
    assert out["extra_keys"] == fwd["extra_keys"]
    U.assert_exact("faces_aug", out["faces_aug"], fwd["faces_aug"])
    U.assert_exact("verts_aug", out["verts_aug"], fwd["verts_aug"])
    U.assert_exact("msdf", out["msdf"], fwd["msdf"])
    U.assert_exact("msdf_watertight", out["msdf_watertight"], fwd["msdf_watertight"])
    U.assert_exact("msdf_boundary", out["msdf_boundary"], fwd["msdf_boundary"])
    if wt:
        assert out["n_verts_watertight"] == fwd["n_verts_watertight"]
        U.assert_exact("faces_watertight", out["faces_watertight"], fwd["faces_watertight"])
        U.assert_exact("vertices_watertight", out["vertices_watertight"], fwd["vertices_watertight"])
    # tangents: per-row bound from the condition of the row's scatter sums (random fields hold near-degenerate faces --
    # exact zeros in sdf -- whose normals are rounding residue: those rows, and only those, are unchecked)
    U.assert_tangents_conditioned("v_tng_aug", out["v_tng_aug"], fwd["v_tng_aug"], O.tangent_condition(fwd),
                                  max_unchecked_frac=0.01 if field != "adv" else 0.2)
    g_pos, g_sdf, g_msdf = O.extract_backward(fwd, grads["g_verts_aug"], grads["g_msdf"], grads["g_vertices_watertight"],
                                              grads["g_msdf_watertight"])
    U.assert_close_normwise("grad_pos", g[0], g_pos, U.GRAD_RTOL)
    U.assert_close_normwise("grad_sdf", g[1], g_sdf, U.GRAD_RTOL)
    if typ == "body":
        assert g[2] is None
    else:
        U.assert_close_normwise("grad_msdf", g[2], g_msdf, U.GRAD_RTOL)


def test_integer_intermediates_match_oracle(dev, edges_mode):
    """classification, case codes, sorted edge keys (interp_v), corner array, counts: bit-exact."""
    from d3human_code_b200 import extract as E
    pos, sdf, msdf, tets = _inputs(24, "adv", seed=3)
    fwd = O.extract_forward(pos, sdf, msdf, tets)
    tt = E.packed_tets(torch.tensor(tets, device=dev), pos.shape[0])
    static = E.static_edges_for(tt, pos.shape[0])
    assert (static is not None) == (edges_mode != "sort")
    assert (static is not None and static[6] is not None) == (edges_mode == "scan")
    if static is not None:   # the static table is the sorted list of all distinct tet edges
        ea = np.minimum(tets[:, [0, 0, 0, 1, 1, 2]], tets[:, [1, 2, 3, 2, 3, 3]]).reshape(-1).astype(np.int64)
        eb = np.maximum(tets[:, [0, 0, 0, 1, 1, 2]], tets[:, [1, 2, 3, 2, 3, 3]]).reshape(-1).astype(np.int64)
        uk = np.unique(ea * pos.shape[0] + eb)
        assert static[2] == uk.shape[0]
        U.assert_exact("edge_ab", static[1].cpu().numpy().astype(np.int64), np.stack([uk // pos.shape[0], uk % pos.shape[0]], 1))
        off = static[0].cpu().numpy()
        assert off[0] == 0 and off[-1] == uk.shape[0] and np.all(np.diff(off) >= 0)
    r = E.forward_raw(torch.tensor(pos, device=dev), torch.tensor(sdf, device=dev), torch.tensor(msdf, device=dev), tt,
                      False, True, static=static)
    c = r.frames[0].counts
    assert c["n_valid_tets"] == fwd["fv"] and c["n_tri_tets"] == fwd["t1"] and c["n_quad_tets"] == fwd["t2"]
    assert c["n_verts"] == fwd["n_verts_watertight"] and c["n_faces_aug"] == fwd["faces_aug"].shape[0]
    assert c["bucket_polys"] == tuple(int(x // k) for x, k in zip(fwd["bucket_counts"], (1, 2, 1, 2, 3, 4)))
    edges = r.tape_edges(0).cpu().numpy()
    U.assert_exact("edge_a", edges[:, 0].astype(np.int64), fwd["edge_a"])
    U.assert_exact("edge_b", edges[:, 1].astype(np.int64), fwd["edge_b"])
    U.assert_exact("corners", r.tape_corners(0).cpu().numpy().astype(np.int64), fwd["corners"])


def test_mapped_host_positions_equal_device_positions(dev):
    """extract.mapped_view: the grid positions stay in pinned host memory and the kernels read the rows they need in place
    (a single call, then a batch of frames that share sdf / msdf): outputs are bit-identical to the same calls on device
    tensors, gradients equal up to the order of the float atomics."""
    from d3human_code_b200 import extract as E
    if dev.type != "cuda":
        pytest.skip("needs pinned, device-mapped host memory")
    pos, sdf, msdf, tets = _inputs(24, "capsule", seed=2)
    _, h = _classes()
    ts, tm = torch.tensor(sdf, device=dev), torch.tensor(msdf, device=dev)
    tt = torch.tensor(tets, device=dev)
    frames = np.stack([pos + np.float32(0.003 * f) for f in range(3)]).astype(np.float32)
    host = torch.from_numpy(frames).pin_memory()
    for _ in range(2):      # (the second round runs on the static tables)
        got = {}
        for kind in ("device", "mapped"):
            tp = (torch.from_numpy(frames).to(dev) if kind == "device" else E.mapped_view(host, dev)).requires_grad_(True)
            assert tp.is_cuda
            v, f, _, _, tng, ex = h(tp[0], ts, tm, tt, "cloth")
            outs = E.extract_frames(tp, ts, tm, tt, types="cloth")
            loss = v.sum() * 0.5 + ex["msdf"].sum() + sum((o[0] * (i + 1)).sum() + o[5]["msdf"].sum() for i, o in enumerate(outs))
            loss.backward()
            got[kind] = [tp.grad, v, f, ex["msdf"]] + [t for o in outs for t in (o[0], o[1], o[5]["msdf"])]
            torch.cuda.synchronize()
        for a, b in zip(got["device"][1:], got["mapped"][1:]):
            assert torch.equal(a.detach().cpu(), b.detach().cpu())
        ga, gb = got["device"][0].cpu().numpy(), got["mapped"][0].cpu().numpy()      # (float atomics: order of additions)
        U.assert_close_normwise("grad_pos", gb, ga, U.GRAD_RTOL)
    with pytest.raises(ValueError):
        E.mapped_view(torch.from_numpy(frames))          # not pinned


def test_smplx_layout_split_extraction(dev):
    """config 3: unstructured (scrambled) grid in the script/get_tet_smpl.py layout, cloth then body on the same sdf."""
    g = grids.smplx_layout_grid(32, dilate=0.15, seed=1)
    pos, tets = g["v"], g["f"]
    sdf, msdf = grids.capsule_garment_field(pos)
    for typ in ("cloth", "body"):
        fwd = O.extract_forward(pos, sdf, msdf, tets, -1 if typ == "body" else 1, True)
        out, _ = _run(dev, pos, sdf[:, None], msdf, tets, "hmSDF_Tets", typ, True)
        U.assert_exact("faces_aug", out["faces_aug"], fwd["faces_aug"])
        U.assert_exact("verts_aug", out["verts_aug"], fwd["verts_aug"])
        U.assert_exact("faces_watertight", out["faces_watertight"], fwd["faces_watertight"])


def test_capacity_regrowth_and_reuse(dev):
    """Same (F,N), surface grows 10x then shrinks: outputs stay exact (capacity overflow is recoverable)."""
    pos, tets = grids.kuhn_grid(24)
    p = pos.astype(np.float64)
    for radius in (0.15, 0.9, 0.3, 0.0):
        sdf = (radius - np.linalg.norm(p, axis=-1)).astype(np.float32)
        msdf = (p[:, 1] + 0.05).astype(np.float32)
        fwd = O.extract_forward(pos, sdf, msdf, tets)
        out, _ = _run(dev, pos, sdf, msdf, tets)
        U.assert_exact("faces_aug", out["faces_aug"], fwd["faces_aug"])
        U.assert_exact("verts_aug", out["verts_aug"], fwd["verts_aug"])


def test_input_validation(dev):
    g, _ = _classes()
    pos = torch.zeros(8, 3, device=dev)
    bad = torch.tensor([[0, 1, 2, 8]], device=dev)
    with pytest.raises(IndexError):
        g(pos, torch.zeros(8, device=dev), torch.zeros(8, device=dev), bad)
    with pytest.raises(ValueError):
        g(pos, torch.zeros(7, device=dev), torch.zeros(8, device=dev), torch.tensor([[0, 1, 2, 3]], device=dev))


def test_msdf_boundary_view_carries_gradient(dev):
    """extra['msdf_boundary'] is msdf[V:] (gshell_tets.py:397): a loss on it must reach the inputs like a loss on the
    slice of extra['msdf'] does in the reference."""
    pos, sdf, msdf, tets = _inputs(16, "capsule")
    fwd = O.extract_forward(pos, sdf, msdf, tets)
    v = fwd["n_verts_watertight"]
    rng = np.random.default_rng(3)
    g_bnd = rng.standard_normal(fwd["msdf"].shape[0] - v).astype(np.float32)
    g_full = rng.standard_normal(fwd["msdf"].shape).astype(np.float32)
    g, _ = _classes()
    tp = torch.tensor(pos, device=dev, requires_grad=True)
    ts = torch.tensor(sdf, device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    verts, faces, _, _, _, extra = g(tp, ts, tm, torch.tensor(tets, device=dev))
    assert extra["msdf_boundary"].data_ptr() == extra["msdf"][v:].data_ptr()
    assert torch.equal(extra["msdf_boundary"], extra["msdf"][v:])
    ((extra["msdf_boundary"] * torch.tensor(g_bnd, device=dev)).sum() + (extra["msdf"] * torch.tensor(g_full, device=dev)).sum()).backward()
    want_up = g_full.copy()
    want_up[v:] += g_bnd
    g_pos, g_sdf, g_msdf = O.extract_backward(fwd, np.zeros_like(fwd["verts_aug"]), want_up)
    U.assert_close_normwise("grad_sdf", ts.grad.cpu().numpy(), g_sdf, U.GRAD_RTOL)
    U.assert_close_normwise("grad_msdf", tm.grad.cpu().numpy(), g_msdf, U.GRAD_RTOL)
    U.assert_close_normwise("grad_pos", tp.grad.cpu().numpy(), g_pos, U.GRAD_RTOL)


def _tangent_upstream(fwd, rng, kappa_max=15.0):
    """Upstream gradients of v_tng_aug / v_tng_watertight that vanish on ill-conditioned rows (where face normals or
    tangents cancel the gradient of the unit vector grows like 1 / |sum| -- in the reference's autograd as well)."""
    kap = O.tangent_condition(fwd)
    nv = fwd["n_verts_watertight"]
    ok = (kap < kappa_max).astype(np.float32)[:, None]
    return (rng.standard_normal(fwd["v_tng_aug"].shape).astype(np.float32) * ok,
            rng.standard_normal((nv, 3)).astype(np.float32) * ok[:nv])


@pytest.mark.parametrize("res,field,cls,typ", [(12, "sphere", "GShell_Tets", None), (16, "capsule", "hmSDF_Tets", "cloth"),
                                               (14, "capsule", "hmSDF_Tets", "body"), (24, "sphere", "hmSDF_Tets", "cloth")])
def test_tangent_gradients_match_oracle(dev, res, field, cls, typ):
    """SURVEY A.5 optional branch: gradients through v_tng (normals, per-face tangents from the vertex-id UVs, Gram-Schmidt,
    boundary interpolation of the tangents) -- d3h_tangent_backward against the oracle's float64 adjoint (pinned against
    the reference's autograd in tests/test_oracle_vs_reference.py), alone and together with the other upstream gradients."""
    pos, sdf, msdf, tets = _inputs(res, field)
    rng = np.random.default_rng(res)
    pos = (pos + 0.2 / res * rng.standard_normal(pos.shape)).astype(np.float32)      # break the lattice symmetry
    fwd = O.extract_forward(pos, sdf, msdf, tets, -1 if typ == "body" else 1, True)
    g_aug, g_wt = _tangent_upstream(fwd, rng)
    g, h = _classes()
    tt = torch.tensor(tets, device=dev)
    for with_rest in (False, True):
        tp = torch.tensor(pos, device=dev, requires_grad=True)
        ts = torch.tensor(sdf, device=dev, requires_grad=True)
        tm = torch.tensor(msdf, device=dev, requires_grad=True)
        verts, faces, _, _, v_tng, extra = g(tp, ts, tm, tt) if cls == "GShell_Tets" else h(tp, ts, tm, tt, typ)
        loss = (v_tng * torch.tensor(g_aug, device=dev)).sum() + (extra["v_tng_watertight"] * torch.tensor(g_wt, device=dev)).sum()
        gv = gm = None
        if with_rest:
            gv = rng.standard_normal(fwd["verts_aug"].shape).astype(np.float32)
            gm = rng.standard_normal(fwd["msdf"].shape).astype(np.float32)
            loss = loss + (verts * torch.tensor(gv, device=dev)).sum() + (extra["msdf"] * torch.tensor(gm, device=dev)).sum()
        loss.backward()
        g_pos, g_sdf, g_msdf = O.extract_backward(fwd, gv, gm, None, None, g_aug, g_wt)
        U.assert_close_normwise("grad_pos", tp.grad.cpu().numpy(), g_pos, 1e-4)
        U.assert_close_normwise("grad_sdf", ts.grad.cpu().numpy(), g_sdf, 1e-4)
        if typ == "body":
            assert tm.grad is None
        else:
            U.assert_close_normwise("grad_msdf", tm.grad.cpu().numpy(), g_msdf, 1e-4)


def test_tangent_gradients_in_a_batch_and_three_face_refusal(dev):
    """The same through extract_frames (one autograd node for several frames; only some frames carry tangent gradients);
    the exactly-three-faces mesh (torch.cross without dim) is refused."""
    from d3human_code_b200.extract import extract_frames
    res = 14
    pos, sdf, msdf, tets = _inputs(res, "capsule")
    rng = np.random.default_rng(3)
    pos_b = np.stack([(pos + 0.2 / res * rng.standard_normal(pos.shape)).astype(np.float32) for _ in range(3)])
    fwds = [O.extract_forward(pos_b[i], sdf, msdf, tets, 1, True) for i in range(3)]
    tp = torch.tensor(pos_b, device=dev, requires_grad=True)
    ts = torch.tensor(sdf, device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    outs = extract_frames(tp, ts, tm, torch.tensor(tets, device=dev), types="cloth", lanes=2)
    ups = [_tangent_upstream(f, rng) for f in fwds]
    gv1 = rng.standard_normal(fwds[1]["verts_aug"].shape).astype(np.float32)
    loss = (outs[0][4] * torch.tensor(ups[0][0], device=dev)).sum() + (outs[1][0] * torch.tensor(gv1, device=dev)).sum() \
        + (outs[2][5]["v_tng_watertight"] * torch.tensor(ups[2][1], device=dev)).sum()
    loss.backward()
    want_sdf = np.zeros_like(sdf, dtype=np.float64)
    wants = [O.extract_backward(fwds[0], None, None, None, None, ups[0][0], None),
             O.extract_backward(fwds[1], gv1, None),
             O.extract_backward(fwds[2], None, None, None, None, None, ups[2][1])]
    for i, w in enumerate(wants):
        U.assert_close_normwise(f"grad_pos[{i}]", tp.grad[i].cpu().numpy(), w[0], 1e-4)
        want_sdf += w[1]
    U.assert_close_normwise("grad_sdf", ts.grad.cpu().numpy(), want_sdf, 1e-4)
    rec = U.load_golden([n for n in U.golden_cases() if "three" in n or "3face" in n or "cross" in n][0]) \
        if [n for n in U.golden_cases() if "three" in n or "3face" in n or "cross" in n] else None
    if rec is not None:
        g, h = _classes()
        t3 = torch.tensor(rec["pos"], device=dev, requires_grad=True)
        out = g(t3, torch.tensor(rec["sdf"], device=dev), torch.tensor(rec["msdf"], device=dev), torch.tensor(rec["tets"], device=dev))
        if out[5]["faces_watertight"].shape[0] == 3:
            with pytest.raises(NotImplementedError):
                out[4].sum().backward()


# ---------------------------------------------------------------------------------------------- batches of frames
def test_extract_frames_batch_matches_oracle_per_frame(dev):
    """BASELINE configs[3]: a batch of frames with per-frame offsets, shared sdf / msdf, mixed cloth / body types, run as
    one autograd node on concurrent lanes: every frame bit-exact against the oracle, shared gradients = sum over frames."""
    from d3human_code_b200.extract import extract_frames, last_counts_frames
    res = 20
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = grids.capsule_garment_field(pos)
    B = 5
    types = ["cloth", "body", "cloth", "cloth", "body"]
    pos_b = np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in range(B)]).astype(np.float32)
    tp = torch.tensor(pos_b, device=dev, requires_grad=True)          # (B,N,3) tensor form
    ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)   # shared, (N,1) like the SDF MLP output
    tm = torch.tensor(msdf, device=dev, requires_grad=True)           # shared
    tt = torch.tensor(tets, device=dev)
    outs = extract_frames(tp, ts, tm, tt, types=types, lanes=3)
    assert len(outs) == B and len(last_counts_frames()) == B
    rng = np.random.default_rng(11)
    want_pos = np.zeros_like(pos_b)
    want_sdf = np.zeros_like(sdf)
    want_msdf = np.zeros_like(msdf)
    loss = 0.0
    for i, (verts, faces, uvs, uv_idx, v_tng, extra) in enumerate(outs):
        fwd = O.extract_forward(pos_b[i], sdf, msdf, tets, -1 if types[i] == "body" else 1, True)
        U.assert_exact(f"faces_aug[{i}]", faces.cpu().numpy(), fwd["faces_aug"])
        U.assert_exact(f"verts_aug[{i}]", verts.detach().cpu().numpy(), fwd["verts_aug"])
        U.assert_exact(f"msdf[{i}]", extra["msdf"].detach().cpu().numpy(), fwd["msdf"])
        U.assert_exact(f"faces_watertight[{i}]", extra["faces_watertight"].cpu().numpy(), fwd["faces_watertight"])
        U.assert_exact(f"vertices_watertight[{i}]", extra["vertices_watertight"].detach().cpu().numpy(),
                       fwd["vertices_watertight"])
        assert extra["n_verts_watertight"] == fwd["n_verts_watertight"]
        gv = rng.standard_normal(fwd["verts_aug"].shape).astype(np.float32)
        gm = rng.standard_normal(fwd["msdf"].shape).astype(np.float32)
        if i == 3:   # one frame receives no upstream gradient at all
            continue
        loss = loss + (verts * torch.tensor(gv, device=dev)).sum() + (extra["msdf"] * torch.tensor(gm, device=dev)).sum()
        g_pos, g_sdf, g_msdf = O.extract_backward(fwd, gv, gm)
        want_pos[i] = g_pos
        want_sdf += g_sdf
        if types[i] != "body":
            want_msdf += g_msdf
    loss.backward()
    U.assert_close_normwise("grad_pos", tp.grad.cpu().numpy(), want_pos, U.GRAD_RTOL)
    U.assert_close_normwise("grad_sdf", ts.grad[:, 0].cpu().numpy(), want_sdf, U.GRAD_RTOL)
    U.assert_close_normwise("grad_msdf", tm.grad.cpu().numpy(), want_msdf, U.GRAD_RTOL)


def test_fused_frames_shared_topology_and_regrowth(dev, edges_mode):
    """Frames of one type on the run-length path run fused (grid.y = frame) and, sharing sdf / msdf, find their topology
    once: 11 frames on 4 workspaces (three rounds of launches) against the oracle frame by frame, gradients of the shared
    tensors summed over the frames -- then once more after the plan's capacities were cut down (every list overflows: the
    call reports it and the host re-runs with room)."""
    from d3human_code_b200 import extract as E
    res, B = 20, 11
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = grids.capsule_garment_field(pos)
    pos_b = np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in range(B)]).astype(np.float32)
    ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    tt = torch.tensor(tets, device=dev)
    rng = np.random.default_rng(5)
    fwd = [O.extract_forward(pos_b[i], sdf, msdf, tets, 1, True) for i in range(B)]
    ups = [(rng.standard_normal(f["verts_aug"].shape).astype(np.float32), rng.standard_normal(f["msdf"].shape).astype(np.float32))
           for f in fwd]
    want_sdf, want_msdf = np.zeros_like(sdf), np.zeros_like(msdf)
    want_pos = np.zeros_like(pos_b)
    for i, f in enumerate(fwd):
        w = O.extract_backward(f, ups[i][0], ups[i][1])
        want_pos[i], want_sdf, want_msdf = w[0], want_sdf + w[1], want_msdf + w[2]
    for shrink in (False, False, True):
        if shrink:
            for plan in E._plans.values():
                plan.cap_tets, plan.cap_v, plan.cap_va, plan.cap_fw, plan.cap_fa = 96, 48, 96, 48, 48
        tp = torch.tensor(pos_b, device=dev, requires_grad=True)
        ts.grad = tm.grad = None
        outs = E.extract_frames(tp, ts, tm, tt, types="cloth", lanes=4)
        loss = 0.0
        for i, o in enumerate(outs):
            U.assert_exact(f"faces_aug[{i}]", o[1].cpu().numpy(), fwd[i]["faces_aug"])
            U.assert_exact(f"verts_aug[{i}]", o[0].detach().cpu().numpy(), fwd[i]["verts_aug"])
            U.assert_exact(f"msdf[{i}]", o[5]["msdf"].detach().cpu().numpy(), fwd[i]["msdf"])
            U.assert_exact(f"faces_watertight[{i}]", o[5]["faces_watertight"].cpu().numpy(), fwd[i]["faces_watertight"])
            loss = loss + (o[0] * torch.tensor(ups[i][0], device=dev)).sum() + (o[5]["msdf"] * torch.tensor(ups[i][1], device=dev)).sum()
        loss.backward()
        U.assert_close_normwise("grad_pos", tp.grad.cpu().numpy(), want_pos, U.GRAD_RTOL)
        U.assert_close_normwise("grad_sdf", ts.grad.cpu().numpy().reshape(-1), want_sdf, 1e-4)
        U.assert_close_normwise("grad_msdf", tm.grad.cpu().numpy(), want_msdf, 1e-4)


def test_extract_frames_list_form_and_repeat(dev):
    """Sequence-of-tensors form with per-frame sdf; two consecutive batches reuse the lanes / graphs / workspaces."""
    from d3human_code_b200.extract import extract_frames
    res = 16
    pos, tets = grids.kuhn_grid(res)
    tt = torch.tensor(tets, device=dev)
    p = pos.astype(np.float64)
    for rep in range(2):
        radii = [0.3 + 0.1 * rep, 0.5, 0.7, 0.0]
        sdfs = [(r - np.linalg.norm(p, axis=-1)).astype(np.float32) for r in radii]
        msdf = (p[:, 1] + 0.05).astype(np.float32)
        tp = torch.tensor(pos, device=dev)
        outs = extract_frames([tp] * 4, [torch.tensor(s, device=dev) for s in sdfs], torch.tensor(msdf, device=dev), tt,
                              lanes=2)
        for i, (verts, faces, _, _, _, extra) in enumerate(outs):
            fwd = O.extract_forward(pos, sdfs[i], msdf, tets)
            U.assert_exact(f"faces_aug[{i}]", faces.cpu().numpy(), fwd["faces_aug"])
            U.assert_exact(f"verts_aug[{i}]", verts.cpu().numpy(), fwd["verts_aug"])


def test_packed_batch_equals_per_frame_results(dev):
    """extract_frames_async(...).packed(): the same batch as padded (B, cap, ...) tensors and ONE autograd node with O(1)
    outputs.  Valid rows equal the per-frame API bit for bit, `frame(i)` narrows to the reference tuple, gradients through
    the padded tensors (upstream rows beyond the valid ones are poison) equal the per-frame gradients."""
    from d3human_code_b200.extract import extract_frames, extract_frames_async
    res, B = 20, 5
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = grids.capsule_garment_field(pos)
    types = ["cloth", "body", "cloth", "cloth", "body"]
    pos_b = np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in range(B)]).astype(np.float32)
    tt = torch.tensor(tets, device=dev)

    def leaves():
        return (torch.tensor(pos_b, device=dev, requires_grad=True), torch.tensor(sdf[:, None], device=dev, requires_grad=True),
                torch.tensor(msdf, device=dev, requires_grad=True))

    tp, ts, tm = leaves()
    outs = extract_frames(tp, ts, tm, tt, types=types, lanes=3)
    rng = np.random.default_rng(2)
    ups = []
    for verts, faces, _, _, _, extra in outs:
        ups.append((torch.tensor(rng.standard_normal(tuple(verts.shape)).astype(np.float32), device=dev),
                    torch.tensor(rng.standard_normal(tuple(extra["msdf"].shape)).astype(np.float32), device=dev),
                    torch.tensor(rng.standard_normal(tuple(extra["vertices_watertight"].shape)).astype(np.float32), device=dev)))
    torch.autograd.backward([o[0] for o in outs] + [o[5]["msdf"] for o in outs] + [o[5]["vertices_watertight"] for o in outs],
                            [u[0] for u in ups] + [u[1] for u in ups] + [u[2] for u in ups])
    for rep in range(2):      # twice: the second batch reuses graphs, workspaces and predicted capacities
        tp2, ts2, tm2 = leaves()
        pk = extract_frames_async(tp2, ts2, tm2, tt, types=types, lanes=3).packed()
        assert len(pk) == B and pk.verts_aug.dim() == 3 and pk.faces_aug.dtype == torch.int64
        gv = torch.full_like(pk.verts_aug, float("nan"))
        gm = torch.full_like(pk.msdf, float("nan"))
        gw = torch.full_like(pk.vertices_watertight, float("nan"))
        for i, (verts, faces, _, _, v_tng, extra) in enumerate(outs):
            va, v, fa, fw = (int(pk.n_verts_aug[i]), int(pk.n_verts_watertight[i]), int(pk.n_faces_aug[i]),
                             int(pk.n_faces_watertight[i]))
            assert (va, fa, v, fw) == (verts.shape[0], faces.shape[0], extra["n_verts_watertight"], extra["faces_watertight"].shape[0])
            assert torch.equal(pk.verts_aug[i, :va].detach(), verts.detach()) and torch.equal(pk.faces_aug[i, :fa], faces)
            assert torch.equal(pk.msdf[i, :va].detach(), extra["msdf"].detach())
            assert torch.equal(pk.vertices_watertight[i, :v].detach(), extra["vertices_watertight"].detach())
            assert torch.equal(pk.faces_watertight[i, :fw], extra["faces_watertight"])
            assert torch.equal(pk.msdf_watertight[i, :v].detach(), extra["msdf_watertight"].detach())
            f = pk.frame(i)
            assert torch.equal(f[0].detach(), verts.detach()) and torch.equal(f[1], faces) and f[2] is None and f[3] is None
            assert tuple(f[5].keys()) == tuple(extra.keys()) and f[5]["n_verts_watertight"] == v
            assert torch.equal(f[5]["msdf_boundary"].detach(), extra["msdf_boundary"].detach())
            gv[i, :va], gm[i, :va], gw[i, :v] = ups[i]
        torch.autograd.backward([pk.verts_aug, pk.msdf, pk.vertices_watertight], [gv, gm, gw])
        U.assert_close_normwise("grad_pos", tp2.grad.cpu().numpy(), tp.grad.cpu().numpy(), U.GRAD_RTOL)
        U.assert_close_normwise("grad_sdf", ts2.grad.cpu().numpy(), ts.grad.cpu().numpy(), U.GRAD_RTOL)
        U.assert_close_normwise("grad_msdf", tm2.grad.cpu().numpy(), tm.grad.cpu().numpy(), U.GRAD_RTOL)


def test_compact_gradient_return(dev):
    """FramesFuture.tape_edges(i) + gather_touched: the dense pos gradient of a frame is zero outside the end points of its
    crossing edges, and the gathered rows restore it exactly (what a host-side consumer copies back instead of (N,3))."""
    from d3human_code_b200.extract import extract_frames_async, gather_touched
    res, B = 20, 3
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = grids.capsule_garment_field(pos)
    pos_b = np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in range(B)]).astype(np.float32)
    tp = torch.tensor(pos_b, device=dev, requires_grad=True)
    ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    fut = extract_frames_async(tp, ts, tm, torch.tensor(tets, device=dev), types="cloth", lanes=2)
    with pytest.raises(RuntimeError):
        fut.tape_edges(0)
    outs = fut.result()
    torch.autograd.backward([o[0].square().sum() + o[5]["msdf"].sum() for o in outs])
    for i, o in enumerate(outs):
        edges = fut.tape_edges(i)
        fwd = O.extract_forward(pos_b[i], sdf, msdf, tets)
        assert edges.dtype == torch.int32 and np.array_equal(edges.cpu().numpy(), np.stack([fwd["edge_a"], fwd["edge_b"]], 1))
        rows = gather_touched(tp.grad[i], edges)
        assert rows.shape == (2 * edges.shape[0], 3)
        dense = torch.zeros_like(tp.grad[i])
        dense[edges.reshape(-1).long()] = rows
        assert torch.equal(dense, tp.grad[i])
    assert gather_touched(tp.grad[0], fut.tape_edges(0)[:0]).shape == (0, 3)
    col = gather_touched(ts.grad, fut.tape_edges(0))      # (N,1) field gradients work the same way
    assert torch.equal(col[:, 0], ts.grad[fut.tape_edges(0).reshape(-1).long(), 0])


# ---------------------------------------------------------------------------------------------- tet-range sharding
@pytest.mark.parametrize("res,field,typ,vr", [(24, "capsule", "cloth", 3), (20, "adv", "body", 2), (33, "sphere", "cloth", 8)])
def test_tet_range_sharding_virtual_ranks_bit_identical(dev, res, field, typ, vr):
    """BASELINE configs[4] path on one GPU: classify `vr` tet ranges separately (d3h_classify_range), concatenate the
    records in range order, run the surface stages on the merged list (d3h_extract_from_records): identical to the
    single call, forward and backward."""
    from d3human_code_b200.extract import extract
    from d3human_code_b200.sharding import extract_tet_sharded
    pos, sdf, msdf, tets = _inputs(res, field, seed=res)
    tt = torch.tensor(tets, device=dev)
    outs = []
    for fn in (lambda *a: extract(*a, msdf_negate=(typ == "body")),
               lambda *a: extract_tet_sharded(*a, msdf_negate=(typ == "body"), virtual_ranks=vr)):
        tp = torch.tensor(pos, device=dev, requires_grad=True)
        ts = torch.tensor(sdf, device=dev, requires_grad=True)
        tm = torch.tensor(msdf, device=dev, requires_grad=True)
        verts, faces, _, _, v_tng, extra = fn(tp, ts, tm, tt)
        (verts.square().sum() + extra["msdf"].sum() + extra["vertices_watertight"].sum()).backward()
        outs.append((verts.detach(), faces, extra["msdf"].detach(), extra["faces_watertight"],
                     extra["vertices_watertight"].detach(), tp.grad, ts.grad))
    for a, b in zip(*outs):
        assert a.shape == b.shape
    for k in range(5):
        assert torch.equal(outs[0][k], outs[1][k]), k
    for k in (5, 6):   # atomics order differs between runs: compare to gradient tolerance
        U.assert_close_normwise(f"grad{k}", outs[1][k].cpu().numpy(), outs[0][k].cpu().numpy(), U.GRAD_RTOL)
    fwd = O.extract_forward(pos, sdf, msdf, tets, -1 if typ == "body" else 1, True)
    U.assert_exact("faces_aug", outs[1][1].cpu().numpy(), fwd["faces_aug"])
    U.assert_exact("verts_aug", outs[1][0].cpu().numpy(), fwd["verts_aug"])


# ---------------------------------------------------------------------------------------------- full size
def test_full_size_properties_128(dev):
    """BASELINE config 2 (128^3 capsules + garment): counts equal the reference's (measured by running the
    reference), the watertight mesh is closed and manifold, integer outputs are deterministic, and the open mesh only
    references rows that are non-zero."""
    pos, tets = grids.kuhn_grid(128)
    sdf, msdf = grids.capsule_garment_field(pos)
    _, h = _classes()
    tp = torch.tensor(pos, device=dev, requires_grad=True)
    ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    tt = torch.tensor(tets, device=dev)
    verts, faces, _, _, v_tng, extra = h(tp, ts, tm, tt, "cloth")
    from d3human_code_b200.extract import last_counts
    c = last_counts()
    assert (c["n_valid_tets"], c["n_verts"], c["n_tri_tets"], c["n_quad_tets"]) == (57740, 38270, 38944, 18796)
    assert (c["n_faces_watertight"], c["n_verts_aug"], c["n_faces_aug"]) == (76536, 230286, 41914)
    fw = extra["faces_watertight"]
    # closed 2-manifold: every undirected edge of the watertight mesh is shared by exactly two faces
    e = torch.cat([fw[:, [0, 1]], fw[:, [1, 2]], fw[:, [2, 0]]], 0)
    e = torch.sort(e, dim=1).values
    key = e[:, 0] * (c["n_verts"] + 1) + e[:, 1]
    _, cnt = torch.unique(key, return_counts=True)
    assert bool((cnt == 2).all())
    # Euler characteristic of a closed genus-0 surface (the capsule union is one blob)
    assert c["n_verts"] - cnt.numel() + fw.shape[0] == 2
    # open mesh: referenced rows are exactly the non-zero rows of verts_aug
    used = torch.zeros(verts.shape[0], dtype=torch.bool, device=dev)
    used[faces.reshape(-1)] = True
    assert int(used.sum()) == 22052
    assert bool((verts.detach()[~used] == 0).all())
    assert bool(torch.isfinite(v_tng).all())
    # determinism of everything but the atomics-ordered tangents; gradients: pos-grad rows sum like the upstream
    (verts.sum() + extra["msdf"].sum()).backward()
    verts2, faces2, _, _, _, extra2 = h(tp.detach(), ts.detach(), tm.detach(), tt, "cloth")
    assert torch.equal(faces, faces2) and torch.equal(verts.detach(), verts2) and torch.equal(fw, extra2["faces_watertight"])
    # linearity property of the adjoint: d(sum verts)/d(pos) sums to the number of used vertices per axis
    # (each used vertex is a convex-ish combination with weights summing to ~1 up to rounding)
    total = tp.grad.sum(0).cpu().numpy()
    assert np.allclose(total, float(used.sum()), rtol=1e-3)
