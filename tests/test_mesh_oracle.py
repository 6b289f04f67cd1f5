"""oracle/mesh_oracle.py against the live reference render/mesh.py (when /root/reference exists) and against the golden
vectors produced from it (tests/golden/mesh_*.npz, oracle/make_golden_mesh.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import mesh_oracle
from oracle.make_golden_mesh import CASES, case_inputs, upstream
from oracle.ref_loader import load_reference_mesh, reference_available
from tests._util import golden_path

# unit normals: the sums are bit-identical (same scatter order as the CPU reference), the final x / sqrt(.) differs by a
# few ulp because ATen vectorises sqrt / div with approximate instructions (same effect as TNG_ATOL in tests/_util.py)
NRM_ATOL = 1e-6


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden(name):
    z = np.load(golden_path(name))
    pos, faces = z["pos"], z["faces"]
    assert np.array_equal(mesh_oracle.mesh_edges(faces), z["edges"])
    nrm = mesh_oracle.auto_normals(pos, faces)
    assert nrm.dtype == np.float32 and np.abs(nrm - z["v_nrm"]).max() <= NRM_ATOL
    g = mesh_oracle.auto_normals_backward(pos, faces, z["g_nrm"])
    scale = max(1.0, float(np.abs(z["g_pos"]).max()))
    assert np.abs(g - z["g_pos"]).max() <= 1e-5 * scale


@pytest.mark.parametrize("name", CASES)
def test_golden_inputs_are_reproducible(name):
    z = np.load(golden_path(name))
    pos, faces = case_inputs(name)
    assert np.array_equal(pos, z["pos"]) and np.array_equal(faces, z["faces"])
    assert np.array_equal(upstream(name, pos.shape[0]), z["g_nrm"])


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_live_reference(seed):
    ref = load_reference_mesh("cpu")
    rng = np.random.default_rng(seed)
    nv = int(rng.integers(4, 80))
    nf = int(rng.integers(1, 200))
    if seed == 5:
        nf = 3
    pos = rng.standard_normal((nv, 3)).astype(np.float32)
    faces = rng.integers(0, nv, size=(nf, 3)).astype(np.int64)
    g = rng.standard_normal((nv, 3)).astype(np.float32)
    p = torch.tensor(pos, requires_grad=True)
    m = ref.Mesh(p, torch.tensor(faces))
    nm = ref.auto_normals(m)
    (nm.v_nrm * torch.tensor(g)).sum().backward()
    assert np.array_equal(mesh_oracle.mesh_edges(faces), m.edges.numpy())
    assert np.abs(mesh_oracle.auto_normals(pos, faces) - nm.v_nrm.detach().numpy()).max() <= NRM_ATOL
    got = mesh_oracle.auto_normals_backward(pos, faces, g)
    want = p.grad.numpy()
    assert np.abs(got - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max()))
