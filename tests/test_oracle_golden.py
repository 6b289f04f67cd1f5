"""The numpy oracle against the committed golden vectors (made from the live reference by oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import gshell_oracle as O
from tests import _util as U


@pytest.mark.parametrize("name", U.golden_cases())
def test_oracle_matches_golden(name):
    rec = U.load_golden(name)
    fwd = O.extract_forward(rec["pos"], rec["sdf"], rec["msdf"], rec["tets"], rec["sign"], rec["wt"])
    U.check_forward_against_golden(fwd, rec)
    if "grad_pos" in rec:
        g_pos, g_sdf, g_msdf = O.extract_backward(
            fwd, rec["g_verts_aug"], rec["g_msdf"], rec.get("g_vertices_watertight"), rec["g_msdf_watertight"])
        U.check_grads_against_golden(g_pos, g_sdf, g_msdf, rec)


def test_golden_set_covers_edge_cases():
    names = U.golden_cases()
    assert len(names) >= 10
    rec = U.load_golden("outside4")
    assert rec["verts_aug"].shape == (0, 3) and rec["faces_aug"].shape == (0, 3)
    rec = U.load_golden("msdfneg6")
    assert rec["faces_aug"].shape == (0, 3) and not rec["verts_aug"].any()
    assert U.load_golden("adv6_body")["tets"].dtype == np.int32
    assert U.load_golden("adv6_body")["sdf"].dtype == np.float64


def test_used_mask_is_a_local_property():
    """Design fact the CUDA kernels rely on: a watertight vertex is referenced by faces_aug iff its mSDF > 0, and a
    boundary vertex iff its own polygon's cut case references it -- so zeroing (gshell_tets.py:423-427) needs no scatter."""
    for name in ("adv6_gshell", "capsule12_cloth", "adv5_open"):
        rec = U.load_golden(name)
        fwd = O.extract_forward(rec["pos"], rec["sdf"], rec["msdf"], rec["tets"], rec["sign"], rec["wt"])
        nv = fwd["n_verts_watertight"]
        assert np.array_equal(fwd["used"][:nv], fwd["msdf_watertight"] > 0)
        mi = fwd["msdf_watertight"][fwd["corners"]] > 0
        mj = fwd["msdf_watertight"][fwd["_nxt"]] > 0
        assert np.array_equal(fwd["used"][nv:], mi != mj)
