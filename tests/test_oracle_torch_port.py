"""oracle/gshell_torch.py (the plain-PyTorch port that bench.py times as "the reference's way on the same GPU") against the
golden vectors of the live reference: integer outputs exact, positions / mSDF 1e-6, tangents and gradients to the oracle
tolerances.  CPU only; on the GPU the same ops run through torch's CUDA kernels."""
import numpy as np
import pytest
import torch

from oracle import gshell_torch as T
from tests import _util as U


@pytest.mark.parametrize("name", U.golden_cases())
def test_torch_port_matches_golden(name):
    rec = U.load_golden(name)
    pos = torch.tensor(rec["pos"], requires_grad=True)
    sdf = torch.tensor(rec["sdf"], requires_grad=True)
    msdf = torch.tensor(rec["msdf"], requires_grad=True)
    tets = torch.tensor(rec["tets"])
    verts, faces, uvs, uv_idx, v_tng, extra = T.extract(pos, sdf, msdf, tets, rec["sign"], rec["wt"])
    assert uvs is None and uv_idx is None and faces.dtype == torch.int64 and verts.dtype == torch.float32
    U.assert_exact("faces_aug", faces.numpy(), rec["faces_aug"])
    for name_, got, want in (("verts_aug", verts, rec["verts_aug"]), ("msdf", extra["msdf"], rec["extra_msdf"]),
                             ("msdf_watertight", extra["msdf_watertight"], rec["extra_msdf_watertight"]),
                             ("msdf_boundary", extra["msdf_boundary"], rec["extra_msdf_boundary"])):
        U.assert_close_normwise(name_, got.detach().numpy(), want, U.POS_RTOL)
    U.assert_tangents_close("v_tng_aug", v_tng.detach().numpy(), rec["v_tng_aug"], 1e-4)
    if rec["wt"]:
        assert extra["n_verts_watertight"] == int(rec["extra_n_verts_watertight"])
        U.assert_exact("faces_watertight", extra["faces_watertight"].numpy(), rec["extra_faces_watertight"])
        U.assert_close_normwise("vertices_watertight", extra["vertices_watertight"].detach().numpy(),
                                rec["extra_vertices_watertight"], U.POS_RTOL)
    else:
        assert set(extra) == {"msdf", "msdf_watertight", "msdf_boundary"}
    if "grad_pos" in rec and verts.shape[0]:
        loss = (verts * torch.tensor(rec["g_verts_aug"])).sum() + (extra["msdf"] * torch.tensor(rec["g_msdf"])).sum()
        loss = loss + (extra["msdf_watertight"] * torch.tensor(rec["g_msdf_watertight"])).sum()
        if rec["wt"]:
            loss = loss + (extra["vertices_watertight"] * torch.tensor(rec["g_vertices_watertight"])).sum()
        loss.backward()
        U.check_grads_against_golden(pos.grad.numpy(), sdf.grad.numpy(), None if msdf.grad is None else msdf.grad.numpy(), rec)
