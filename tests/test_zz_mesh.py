"""Parity of the mesh stage (d3human_code_b200.render.mesh -> C ABI of include/d3h_mesh.h -> sm_100a
kernels) with the golden vectors produced by the live reference `render/mesh.py`, with the numpy oracle on seeded
meshes, and through size-independent properties on a full-size (128^3) extraction.

Tolerances: edges bit-exact (values and order); unit normals 2e-6 + 4e-7 x condition absolute, where the condition of a
vertex is sum|face normal| / |sum of face normals| (float atomics add in another order than the CPU reference, which
moves the result by ~1.3e-7 x condition -- measured by re-ordering the faces in the oracle; the reference's own CUDA
scatter_add has the same freedom); vertices whose face normals cancel (condition >= 1e4, e.g. two in the coarse sphere
fixture) have no stable normal or gradient and are only required to come out unit length; gradients 1e-5 normwise.  The file sorts after every other gpu test on purpose: it was written
when no GPU was available (validated on the kernel emulation, tests/test_emu_mesh.py).
"""
import numpy as np
import pytest
import torch

from oracle import gshell_oracle as O
from oracle import mesh_oracle as MO
from oracle.make_golden_mesh import CASES
from d3human_code_b200 import grids
from tests import _util as U

pytestmark = pytest.mark.gpu

NRM_ATOL = 2e-6
NRM_COND = 4e-7
COND_MAX = 1e4        # above: the face normals of the vertex cancel, no stable normal
GRAD_COND_MAX = 1e3   # gradients are compared on meshes whose worst vertex is below this (the adjoint carries 1 / |sum|)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _mesh():
    from d3human_code_b200.render import mesh
    return mesh


def _run(dev, pos, faces, g_nrm=None):
    mesh = _mesh()
    p = torch.tensor(pos, device=dev, requires_grad=True)
    f = torch.tensor(faces, device=dev)
    m = mesh.Mesh(p, f)
    out = {"edges": m.edges.cpu().numpy()}
    nm = mesh.auto_normals(m)
    assert nm.v_nrm.dtype == torch.float32 and nm.t_nrm_idx is f and nm.v_pos is p
    out["v_nrm"] = nm.v_nrm.detach().cpu().numpy()
    if g_nrm is not None:
        (nm.v_nrm * torch.tensor(g_nrm, device=dev)).sum().backward()
        out["g_pos"] = p.grad.cpu().numpy()
    return out


def _check_normals(got, want, cond, what=""):
    stable = cond < COND_MAX
    err = np.abs(got - want).max(1, initial=0.0)
    tol = NRM_ATOL + NRM_COND * cond
    assert bool((err[stable] <= tol[stable]).all()), (what, float(err[stable].max()), int(np.argmax(err * stable)))
    assert np.abs(np.linalg.norm(got.astype(np.float64), axis=1) - 1.0).max(initial=0.0) <= 1e-5, what


def _check(out, edges, v_nrm, cond, g_pos=None, what=""):
    assert out["edges"].dtype == np.int64 and out["edges"].shape == edges.shape, (what, out["edges"].shape, edges.shape)
    assert np.array_equal(out["edges"], edges), what
    _check_normals(out["v_nrm"], v_nrm, cond, what)
    if g_pos is not None:
        assert np.isfinite(out["g_pos"]).all(), what
        if cond.max() < GRAD_COND_MAX:
            U.assert_close_normwise(what + " g_pos", out["g_pos"], g_pos, U.GRAD_RTOL)


@pytest.mark.parametrize("name", CASES)
def test_mesh_matches_golden(dev, name):
    z = np.load(U.golden_path(name))
    if z["faces"].shape[0] == 0:  # the reference cannot build an empty mesh; behaviour fixed by the oracle
        mesh = _mesh()
        p = torch.tensor(z["pos"], device=dev, requires_grad=True)
        m = mesh.Mesh(p, torch.tensor(z["faces"], device=dev))
        assert tuple(m.edges.shape) == (0, 2) and m.edges.dtype == torch.int64
        nm = mesh.auto_normals(m)
        assert np.array_equal(nm.v_nrm.detach().cpu().numpy(), z["v_nrm"])
        nm.v_nrm.sum().backward()
        assert float(p.grad.abs().max()) == 0.0
        return
    out = _run(dev, z["pos"], z["faces"], z["g_nrm"])
    _check(out, z["edges"], z["v_nrm"], MO.normal_condition(z["pos"], z["faces"]), z["g_pos"], name)


def _height_field(n, seed, permute=True):
    """n x n triangulated height field (V = n^2 > one scan tile for n >= 65), vertex labels and face order shuffled."""
    rng = np.random.default_rng(seed)
    x, y = np.meshgrid(np.linspace(-1, 1, n), np.linspace(-1, 1, n), indexing="ij")
    pos = np.stack([x, y, 0.3 * np.sin(3 * x) * np.cos(2 * y)], -1).reshape(-1, 3)
    pos = (pos + 0.2 / n * rng.standard_normal(pos.shape)).astype(np.float32)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a = (i * n + j).reshape(-1)
    faces = np.concatenate([np.stack([a, a + n, a + 1], 1), np.stack([a + 1, a + n, a + n + 1], 1)], 0).astype(np.int64)
    if permute:
        perm = rng.permutation(n * n)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(n * n)
        pos = pos[perm]
        faces = inv[faces][rng.permutation(faces.shape[0])]
    return pos, faces


@pytest.mark.parametrize("n,seed", [(9, 0), (40, 1), (80, 2)])
def test_mesh_matches_oracle(dev, n, seed):
    pos, faces = _height_field(n, seed)
    g = np.random.default_rng(seed + 100).standard_normal(pos.shape).astype(np.float32)
    out = _run(dev, pos, faces, g)
    _check(out, MO.mesh_edges(faces), MO.auto_normals(pos, faces), MO.normal_condition(pos, faces),
           MO.auto_normals_backward(pos, faces, g), f"field{n}")
    assert out["edges"].shape[0] == 3 * (n - 1) * (n - 1) + 2 * (n - 1)  # edges of a triangulated n x n patch


def test_hub_vertex_with_many_neighbours(dev):
    """One vertex adjacent to 700 others and smaller than all of them: the long-segment path of the edge kernels."""
    n = 700
    rng = np.random.default_rng(5)
    pos = rng.standard_normal((n + 1, 3)).astype(np.float32)
    ring = 1 + rng.permutation(n)
    faces = np.stack([ring, np.zeros(n, np.int64), np.roll(ring, 1)], 1).astype(np.int64)
    out = _run(dev, pos, faces)
    _check(out, MO.mesh_edges(faces), MO.auto_normals(pos, faces), MO.normal_condition(pos, faces), None, "hub")


def test_drop_in_semantics(dev):
    mesh = _mesh()
    pos, faces = _height_field(12, 3)
    p = torch.tensor(pos, device=dev)
    f = torch.tensor(faces, device=dev)
    m = mesh.Mesh(p, f, edges="ignored like in the reference")          # get_edge() overwrites it, render/mesh.py:162
    want = MO.mesh_edges(faces)
    assert np.array_equal(m.edges.cpu().numpy(), want)
    assert np.array_equal(m.get_edge().cpu().numpy(), want)
    m2 = mesh.Mesh(p * 2.0, base=m)                                     # hmsdf.py:472: same faces, other positions
    assert m2.t_pos_idx is f and m2.edges is m.edges                    # one edge list per faces tensor
    nm = mesh.auto_normals(m2)
    assert nm.v_pos is m2.v_pos and nm.t_pos_idx is f and nm.t_nrm_idx is f and nm.material is None
    c = nm.clone()
    assert c.v_nrm is not nm.v_nrm and torch.equal(c.v_nrm, nm.v_nrm) and not c.v_nrm.requires_grad
    assert c.t_pos_idx is not f and np.array_equal(c.edges.cpu().numpy(), want)
    with pytest.raises(TypeError):
        mesh.Mesh(p)                                                    # the reference fails in get_edge as well
    if dev.type == "cuda":
        with pytest.raises(RuntimeError):                               # no CPU path
            mesh.Mesh(torch.tensor(pos), torch.tensor(faces)).edges
        with pytest.raises(RuntimeError):
            mesh.auto_normals(mesh.Mesh(torch.tensor(pos), torch.tensor(faces)))
    bad = f.clone()
    bad[3, 1] = pos.shape[0]
    with pytest.raises(IndexError):
        mesh.Mesh(p, bad).edges
    only_faces = mesh.Mesh(t_pos_idx=f)                                 # no positions: index range from the faces
    assert np.array_equal(only_faces.edges.cpu().numpy(), want)
    i32 = mesh.Mesh(p, f.to(torch.int32))
    assert i32.edges.dtype == torch.int64 and np.array_equal(i32.edges.cpu().numpy(), want)
    p64 = torch.tensor(pos, device=dev, dtype=torch.float64, requires_grad=True)
    n64 = mesh.auto_normals(mesh.Mesh(p64, f)).v_nrm
    assert n64.dtype == torch.float64
    n64.sum().backward()
    assert p64.grad.dtype == torch.float64 and bool(torch.isfinite(p64.grad).all())


def test_repeated_auto_normals_share_one_result(dev):
    """hmsdf.py:558-561 / :589-593: the same (verts, faces) goes through auto_normals twice per iteration."""
    mesh = _mesh()
    pos, faces = _height_field(10, 4)
    p = torch.tensor(pos, device=dev, requires_grad=True)
    f = torch.tensor(faces, device=dev)
    before = mesh.launch_counter()
    a = mesh.auto_normals(mesh.Mesh(p, f))
    mid = mesh.launch_counter()
    b = mesh.auto_normals(mesh.Mesh(p, f, material="m"))
    assert b.v_nrm is a.v_nrm and mesh.launch_counter() == mid > before
    with torch.no_grad():
        c = mesh.auto_normals(mesh.Mesh(p, f))
    assert c.v_nrm is not a.v_nrm and not c.v_nrm.requires_grad
    assert float((c.v_nrm - a.v_nrm.detach()).abs().max()) <= NRM_ATOL          # float atomics: not bit-reproducible run to run
    (a.v_nrm.sum() + 2.0 * b.v_nrm.sum()).backward()                   # both users' gradients arrive
    want = MO.auto_normals_backward(pos, faces, np.full(pos.shape, 3.0, np.float32))
    U.assert_close_normwise("g_pos", p.grad.cpu().numpy(), want, U.GRAD_RTOL)
    d = mesh.auto_normals(mesh.Mesh(p.detach().clone().requires_grad_(True), f))
    assert d.v_nrm is not a.v_nrm                                       # another positions tensor: computed again
    with torch.no_grad():
        p.mul_(1.5)                                                     # in-place change of the positions: stale entry
    e = mesh.auto_normals(mesh.Mesh(p, f))
    assert e.v_nrm is not a.v_nrm


def test_mesh_on_extraction_output(dev):
    """The call sequence of hmsdf.py:548-593: extraction -> Mesh -> auto_normals, gradient back to the grid."""
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    mesh = _mesh()
    pos, tets = grids.kuhn_grid(16)
    sdf, msdf = grids.capsule_garment_field(pos)
    tp = torch.tensor(pos, device=dev, requires_grad=True)
    ts = torch.tensor(sdf, device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    verts, faces, _, _, _, extra = hmSDF_Tets()(tp, ts, tm, torch.tensor(tets, device=dev), "cloth")
    im = mesh.auto_normals(mesh.Mesh(verts, faces))
    wt = mesh.auto_normals(mesh.Mesh(extra["vertices_watertight"], extra["faces_watertight"]))
    fwd = O.extract_forward(pos, sdf, msdf, tets)
    assert np.array_equal(im.edges.cpu().numpy(), MO.mesh_edges(fwd["faces_aug"]))
    assert np.array_equal(wt.edges.cpu().numpy(), MO.mesh_edges(fwd["faces_watertight"]))
    cond = MO.normal_condition(fwd["verts_aug"], fwd["faces_aug"])
    assert cond.max() < COND_MAX
    _check_normals(im.v_nrm.detach().cpu().numpy(), MO.auto_normals(fwd["verts_aug"], fwd["faces_aug"]), cond, "open")
    _check_normals(wt.v_nrm.detach().cpu().numpy(), MO.auto_normals(fwd["vertices_watertight"], fwd["faces_watertight"]),
                   MO.normal_condition(fwd["vertices_watertight"], fwd["faces_watertight"]), "watertight")
    g = np.random.default_rng(9).standard_normal(fwd["verts_aug"].shape).astype(np.float32)
    verts.retain_grad()
    (im.v_nrm * torch.tensor(g, device=dev)).sum().backward()
    g_verts = verts.grad.cpu().numpy()
    U.assert_close_normwise("g_verts", g_verts, MO.auto_normals_backward(fwd["verts_aug"], fwd["faces_aug"], g), U.GRAD_RTOL)
    # the extraction's adjoint is checked with the very upstream gradient it received (it amplifies differences in it)
    g_pos, g_sdf, g_m = O.extract_backward(fwd, g_verts_aug=g_verts)
    U.assert_close_normwise("g_pos", tp.grad.cpu().numpy(), g_pos, U.GRAD_RTOL)
    # this upstream gradient spans four orders of magnitude (1 / area factors of the normalisation): the fp32 adjoint of
    # the crossing weights (1 / (sa - sb)^2 terms) sits at 1.2e-5 of the float64 oracle, g_pos / g_msdf stay within 1e-5
    U.assert_close_normwise("g_sdf", ts.grad.cpu().numpy().reshape(-1), g_sdf, 5 * U.GRAD_RTOL)
    U.assert_close_normwise("g_msdf", tm.grad.cpu().numpy(), g_m, U.GRAD_RTOL)


def test_full_size_128(dev):
    """128^3 sphere (BASELINE.json configs[1] size).  The watertight surface is a closed 2-manifold of genus 0, so
    E = 3F/2 and V - E + F = 2; edge lists and normals of the watertight and of the open surface equal the oracle's.
    (The Kuhn lattice mixes left- and right-handed tets, so the reference's triangle table orients neighbouring faces
    inconsistently and thousands of vertex normals nearly cancel: the comparison is condition-aware, see the header.)"""
    if dev.type != "cuda":
        pytest.skip("full size: GPU only")
    from d3human_code_b200.geometry.gshell_tets import GShell_Tets
    mesh = _mesh()
    pos, tets = grids.kuhn_grid(128)
    sdf, msdf = grids.sphere_plane_field(pos)
    verts, faces, _, _, _, extra = GShell_Tets()(torch.tensor(pos, device=dev), torch.tensor(sdf, device=dev),
                                                 torch.tensor(msdf, device=dev), torch.tensor(tets, device=dev))
    fwd = O.extract_forward(pos, sdf, msdf, tets, n_threads=8)
    wt = mesh.Mesh(extra["vertices_watertight"], extra["faces_watertight"])
    V, F, E = wt.v_pos.shape[0], wt.t_pos_idx.shape[0], wt.edges.shape[0]
    assert (V, F) == (83222, 166440)                                    # SURVEY 8(a) [probe]
    assert 2 * E == 3 * F and V - E + F == 2
    e = wt.edges
    assert bool((e[:, 0] < e[:, 1]).all())
    key = e[:, 0] * V + e[:, 1]
    assert bool((key[1:] > key[:-1]).all())                             # strictly ascending lexicographic order
    assert np.array_equal(e.cpu().numpy(), MO.mesh_edges(fwd["faces_watertight"]))
    _check_normals(mesh.auto_normals(wt).v_nrm.cpu().numpy(),
                   MO.auto_normals(fwd["vertices_watertight"], fwd["faces_watertight"]),
                   MO.normal_condition(fwd["vertices_watertight"], fwd["faces_watertight"]), "watertight 128")
    open_mesh = mesh.Mesh(verts, faces)                                 # open surface cut by the mSDF: most rows unused
    n_open = mesh.auto_normals(open_mesh).v_nrm
    used = torch.zeros(verts.shape[0], dtype=torch.bool, device=dev)
    used[faces.reshape(-1)] = True
    assert bool((n_open[~used] == torch.tensor([0.0, 0.0, 1.0], device=dev)).all())
    assert np.array_equal(open_mesh.edges.cpu().numpy(), MO.mesh_edges(fwd["faces_aug"]))
    _check_normals(n_open.cpu().numpy(), MO.auto_normals(fwd["verts_aug"], fwd["faces_aug"]),
                   MO.normal_condition(fwd["verts_aug"], fwd["faces_aug"]), "open 128")
    fe = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]).sort(dim=1).values
    ko = open_mesh.edges[:, 0] * verts.shape[0] + open_mesh.edges[:, 1]
    assert torch.equal(torch.unique(fe[:, 0] * verts.shape[0] + fe[:, 1]), ko)   # same set as a plain sort + unique


@pytest.mark.parametrize("seed", range(12))
def test_random_triangle_soups(dev, seed):
    """Seeded soups: repeated faces and edges, degenerate faces, isolated vertices, dense graphs (every pair of a small
    vertex set is an edge: long hash probes, long neighbour segments)."""
    rng = np.random.default_rng(100 + seed)
    nv = int(rng.integers(3, 400))
    nf = int(rng.integers(1, 6000 if seed % 3 == 0 else 300))
    pos = rng.standard_normal((nv + int(rng.integers(0, 50)), 3)).astype(np.float32)
    faces = rng.integers(0, nv, size=(nf, 3)).astype(np.int64)
    if nf > 4:
        faces[1] = faces[0]
        faces[2] = faces[0][::-1]
        faces[3, 2] = faces[3, 0]
    g = rng.standard_normal(pos.shape).astype(np.float32)
    out = _run(dev, pos, faces, g)
    cond = MO.normal_condition(pos, faces)
    stable_mesh = cond.max() < GRAD_COND_MAX
    _check(out, MO.mesh_edges(faces), MO.auto_normals(pos, faces), cond,
           MO.auto_normals_backward(pos, faces, g) if stable_mesh else None, f"soup{seed}")


@pytest.mark.parametrize("nv", [4095, 4096, 4097, 8192, 12289])
def test_vertex_counts_around_the_scan_tile(dev, nv):
    """Vertex counts at and around multiples of the 4096-vertex scan tile; the last vertices carry edges."""
    rng = np.random.default_rng(nv)
    pos = rng.standard_normal((nv, 3)).astype(np.float32)
    faces = rng.integers(0, nv, size=(3000, 3)).astype(np.int64)
    faces[:8] = np.array([[nv - 1, nv - 2, nv - 3], [nv - 1, 0, nv - 2], [4095 % nv, 4096 % nv, 4094 % nv], [0, 1, nv - 1],
                          [nv - 3, nv - 1, 1], [4096 % nv, 0, nv - 1], [2, 4095 % nv, 4096 % nv], [nv - 2, 4096 % nv, 3]])
    out = _run(dev, pos, faces)
    _check(out, MO.mesh_edges(faces), MO.auto_normals(pos, faces), MO.normal_condition(pos, faces), None, f"nv{nv}")


@pytest.mark.parametrize("seed", range(16))
def test_degenerate_soups(dev, seed):
    """Soups over a handful of vertices: most faces repeat an index (zero normal up to the rounding residue of the fused
    cross product), many vertices are isolated, positions on a coarse lattice make sums cancel exactly.  Edges must be
    exact; normals are compared where they are stable, gradients where no vertex amplifies a residue (see
    oracle/mesh_oracle.normal_condition) -- found by a 400-seed campaign on the kernel emulation."""
    rng = np.random.default_rng(10_000 + seed)
    nv = int(rng.integers(1, 9000)) if seed % 7 == 0 else int(rng.integers(1, 300))
    nf = int(rng.integers(0, 4000)) if seed % 5 == 0 else int(rng.integers(0, 120))
    used = int(rng.integers(1, nv + 1)) if seed % 2 else int(rng.integers(1, 6))
    pos = rng.standard_normal((nv, 3)).astype(np.float32)
    if seed % 3 == 0:
        pos = (np.round(pos * 2) / 2).astype(np.float32)
    faces = rng.integers(0, used, size=(nf, 3)).astype(np.int64)
    g = rng.standard_normal(pos.shape).astype(np.float32)
    out = _run(dev, pos, faces, g if nf else None)
    cond = MO.normal_condition(pos, faces)
    _check(out, MO.mesh_edges(faces), MO.auto_normals(pos, faces), cond,
           MO.auto_normals_backward(pos, faces, g) if nf and cond.max() < GRAD_COND_MAX else None, f"degenerate{seed}")
