"""tests/test_zz_mesh.py (the `-m gpu` parity tests of the mesh stage) run on the CPU against the functional emulation
of the CUDA kernels (tests/emu/): checks the LOGIC of csrc/d3h_mesh.cu and the host path through the real C ABI.  Races,
memory ordering and speed are what the GPU runs are for.  Test infrastructure only."""
import contextlib

import pytest
import torch

from d3human_code_b200 import _cabi
from d3human_code_b200 import extract as E
from d3human_code_b200.render import mesh as M
from tests import test_zz_mesh as Z
from tests.test_emu_parity import _Stream, _packed_cache, _packed_tets_cpu, build_emu


@pytest.fixture(scope="module")
def emu_lib_path():
    return build_emu.build()


@pytest.fixture
def dev(emu_lib_path, monkeypatch):
    monkeypatch.setattr(_cabi, "LIB_PATH", emu_lib_path)
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(M, "_check_cuda", lambda t: None)
    monkeypatch.setattr(E, "_check_cuda", lambda t: None)
    monkeypatch.setattr(E, "packed_tets", _packed_tets_cpu)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: _Stream())
    monkeypatch.setattr(torch.cuda, "device", lambda dev=None: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    orig_ensure = E._Plan.ensure

    def ensure(self, lanes, n_edges=0):       # CPU allocations are 64-byte aligned, the ABI wants 256 for the workspace
        orig_ensure(self, lanes, n_edges)
        for i, w in enumerate(self.workspaces):
            if w.data_ptr() % 256:
                big = torch.empty(self.workspace_bytes + 256, dtype=torch.uint8)
                off = (-big.data_ptr()) % 256
                self.workspaces[i] = big[off:off + self.workspace_bytes]
                self.workspace_ptrs[i] = self.workspaces[i].data_ptr()

    monkeypatch.setattr(E._Plan, "ensure", ensure)
    E.reset_plans()
    M.reset()
    _packed_cache.clear()
    yield torch.device("cpu")
    E.reset_plans()
    M.reset()


@pytest.mark.parametrize("name", Z.CASES)
def test_golden(dev, name):
    Z.test_mesh_matches_golden(dev, name)


@pytest.mark.parametrize("n,seed", [(9, 0), (40, 1), (80, 2)])
def test_oracle(dev, n, seed):
    Z.test_mesh_matches_oracle(dev, n, seed)


def test_hub(dev):
    Z.test_hub_vertex_with_many_neighbours(dev)


def test_drop_in(dev):
    Z.test_drop_in_semantics(dev)


def test_shared_normals(dev):
    Z.test_repeated_auto_normals_share_one_result(dev)


def test_on_extraction_output(dev):
    Z.test_mesh_on_extraction_output(dev)


@pytest.mark.parametrize("seed", range(12))
def test_soups(dev, seed):
    Z.test_random_triangle_soups(dev, seed)


@pytest.mark.parametrize("nv", [4095, 4096, 4097, 8192, 12289])
def test_scan_tile_boundaries(dev, nv):
    Z.test_vertex_counts_around_the_scan_tile(dev, nv)


def test_accelerate_the_reference_module(dev):
    """INTEGRATION.md section 4, second variant: the reference's own render/mesh.py with `Mesh, auto_normals =
    accelerate(Mesh)` at its bottom -- every other function of the module keeps working on the subclass.  Needs the live
    reference (build container only); runs on the emulated kernels."""
    import numpy as np
    from oracle.ref_loader import load_reference_mesh, reference_available
    if not reference_available():
        pytest.skip("reference tree not present")
    ref = load_reference_mesh("cpu")
    ref_mesh_cls, ref_auto_normals = ref.Mesh, ref.auto_normals
    pos, faces = Z._height_field(14, 7)
    p = torch.tensor(pos, requires_grad=True)
    f = torch.tensor(faces)
    want_mesh = ref_auto_normals(ref_mesh_cls(p, f))                    # the reference, untouched
    (want_mesh.v_nrm * 2.0).sum().backward()
    want_grad, p.grad = p.grad.clone(), None

    ref.Mesh, ref.auto_normals = M.accelerate(ref.Mesh)                 # what the maintainer adds
    before = M.launch_counter()
    m = ref.Mesh(p, f, material="mat")
    assert isinstance(m, ref_mesh_cls) and M.launch_counter() == before  # no edge computation in the constructor
    assert np.array_equal(m.edges.numpy(), want_mesh.edges.numpy()) and M.launch_counter() > before
    nm = ref.auto_normals(m)
    assert isinstance(nm, ref.Mesh) and nm.material == "mat" and nm.t_nrm_idx is f
    assert float((nm.v_nrm - want_mesh.v_nrm).detach().abs().max()) <= Z.NRM_ATOL
    (nm.v_nrm * 2.0).sum().backward()
    assert float((p.grad - want_grad).abs().max()) <= 1e-5 * float(want_grad.abs().max())
    c = nm.clone()                                                      # the reference's own clone(), :203-238
    assert isinstance(c, ref.Mesh) and torch.equal(c.v_nrm, nm.v_nrm.detach()) and torch.equal(c.edges, m.edges)
    lo, hi = ref.aabb(nm)                                               # an untouched helper of the reference module
    assert torch.equal(lo, p.detach().min(0).values) and torch.equal(hi, p.detach().max(0).values)
    n2 = ref.unit_size(nm)                                              # builds a Mesh through the module's global name
    assert isinstance(n2, ref.Mesh) and np.array_equal(n2.edges.numpy(), want_mesh.edges.numpy())


@pytest.mark.parametrize("seed", range(16))
def test_degenerate_soups(dev, seed):
    Z.test_degenerate_soups(dev, seed)
