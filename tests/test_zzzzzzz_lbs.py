"""Skinning of the extracted vertices (include/d3h_lbs.h, d3human-code_b200/deform/lbs.py) against the
float64 oracle (oracle/lbs_oracle.py) and the golden vectors written from the reference's own methods
(tests/golden/lbs_*.npz).  GPU only.  Tolerances: positions 2e-5 (the reference inverts the blended 4x4 per point in
fp32, the oracle in float64), gradients 1e-4 normwise."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import lbs_oracle as LO
from tests.test_lbs_oracle import synthetic_rig

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lbs_*.npz")))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _rig(dev, template, w, init_a):
    from d3human_code_b200.deform.lbs import LinearBlendSkinning
    return LinearBlendSkinning(torch.tensor(template[None], device=dev), torch.tensor(w, device=dev), torch.tensor(init_a[None], device=dev))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_lbs_matches_golden_vectors_of_the_reference(dev, path):
    d = np.load(path)
    rig = _rig(dev, d["template"], d["w"], d["init_a"])
    tp = torch.tensor(d["pts"][None], device=dev, requires_grad=True)
    ta = torch.tensor(d["a"][None], device=dev, requires_grad=True)
    tt = torch.tensor(d["trans"].reshape(1, 3), device=dev, requires_grad=True)
    out = rig.lbs_transform(tp, ta, tt)
    assert out.shape == (d["pts"].shape[0], 3)
    assert np.abs(out.detach().cpu().numpy() - d["posed"]).max() < 2e-5 * max(1.0, np.abs(d["posed"]).max())
    assert np.abs(rig.lbs_forward_inverse(tp).cpu().numpy()[0] - d["canonical"]).max() < 2e-5
    (out * torch.tensor(d["g"], device=dev)).sum().backward()
    for name, got, want in (("pts", tp.grad[0], d["g_pts"]), ("A", ta.grad[0], d["g_a"]), ("trans", tt.grad[0], d["g_trans"])):
        err = np.abs(got.cpu().numpy() - want).max() / np.abs(want).max()
        assert err < 1e-4, (name, err)


@pytest.mark.parametrize("seed,p,vt,j,zero_frac", [(0, 1, 50, 4, 0.0), (1, 1000, 300, 8, 0.8), (2, 33333, 10475, 55, 0.8),
                                                   (3, 5000, 2000, 55, 1.0)])
def test_lbs_matches_oracle(dev, seed, p, vt, j, zero_frac):
    """Random rigs up to the SMPL-X sizes (10 475 template vertices, 55 joints), with the share of exactly-zero rows
    verts_aug has (gshell_tets.py:423-427): nearest-vertex indices exact, positions 2e-5, gradients 1e-4."""
    template, w, init_a, a, trans = synthetic_rig(seed, vt, j)
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1.1, 1.1, size=(p, 3)).astype(np.float32)
    pts[rng.random(p) < zero_frac] = 0.0
    rig = _rig(dev, template, w, init_a)
    idx = rig.nearest(torch.tensor(pts, device=dev)).cpu().numpy()
    want_idx = LO.nearest(pts, template)
    bad = idx != want_idx            # fp32 vs float64 distances may order a near-tie differently: the distances must agree
    if bad.any():
        d_got = ((pts[bad] - template[idx[bad]]) ** 2).sum(-1)
        d_want = ((pts[bad] - template[want_idx[bad]]) ** 2).sum(-1)
        assert np.all(np.abs(d_got - d_want) <= 1e-6 * (1 + d_want)) and bad.mean() < 1e-3
    assert np.array_equal(rig.interpolate_weights(torch.tensor(pts[None], device=dev)).cpu().numpy()[0], w[idx])
    tp = torch.tensor(pts, device=dev, requires_grad=True)
    ta = torch.tensor(a, device=dev, requires_grad=True)
    tt = torch.tensor(trans, device=dev, requires_grad=True)
    out = rig.lbs_transform(tp, ta, tt)
    ok = ~bad
    want, cache = LO.lbs_forward(pts, template, w, init_a, a, trans)
    assert np.abs(out.detach().cpu().numpy() - want)[ok].max() < 2e-5 * max(1.0, np.abs(want).max())
    if bad.any():
        return
    g = rng.standard_normal(pts.shape).astype(np.float32)
    (out * torch.tensor(g, device=dev)).sum().backward()
    o_pts, o_a, o_t = LO.lbs_backward(cache, g)
    for name, got, w_ in (("pts", tp.grad, o_pts), ("A", ta.grad, o_a), ("trans", tt.grad, o_t)):
        err = np.abs(got.cpu().numpy() - w_).max() / max(np.abs(w_).max(), 1e-30)
        assert err < 1e-4, (name, err)


def test_lbs_on_an_extracted_surface_and_k_restriction(dev):
    """The stage as D3-Human uses it (hmsdf.py:471): verts_aug of an extraction (mostly zero rows) through the skinning,
    gradient back through the extraction to the grid."""
    from d3human_code_b200 import grids
    from d3human_code_b200.deform.lbs import LinearBlendSkinning
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    pos, tets = grids.kuhn_grid(20)
    sdf, msdf = grids.capsule_garment_field(pos)
    template, w, init_a, a, trans = synthetic_rig(7, 3000, 55)
    rig = _rig(dev, template, w, init_a)
    tp = torch.tensor(pos, device=dev, requires_grad=True)
    ts = torch.tensor(sdf, device=dev, requires_grad=True)
    verts, faces, *_ = hmSDF_Tets()(tp, ts, torch.tensor(msdf, device=dev), torch.tensor(tets, device=dev), "cloth")
    assert float((verts.detach().abs().sum(1) == 0).float().mean()) > 0.5
    out = rig.lbs_transform(verts.reshape(1, -1, 3), torch.tensor(a[None], device=dev), torch.tensor(trans.reshape(1, 3), device=dev))
    want, _ = LO.lbs_forward(verts.detach().cpu().numpy(), template, w, init_a, a, trans)
    assert np.abs(out.detach().cpu().numpy() - want).max() < 3e-5 * max(1.0, np.abs(want).max())
    out.square().sum().backward()
    assert ts.grad is not None and torch.isfinite(ts.grad).all() and float(ts.grad.abs().max()) > 0
    with pytest.raises(NotImplementedError):
        LinearBlendSkinning(torch.tensor(template, device=dev), torch.tensor(w, device=dev), torch.tensor(init_a, device=dev), k=4)
