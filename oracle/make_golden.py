"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

    python oracle/make_golden.py

Imports the reference classes from /root/reference through oracle/ref_loader.py, runs them on CPU
(`GShell_Tets.__call__` gshell_tets.py:253, `hmSDF_Tets.__call__` hmsdf_tets_split.py:254) on small seeded
inputs and stores inputs, every returned tensor and the autograd gradients for seeded upstream gradients.
The fixtures pin the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_cuda_parity.py) on
machines where the reference tree does not exist.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference_class  # noqa: E402
from d3human_code_b200 import grids  # noqa: E402

warnings.filterwarnings("ignore")
OUT = os.path.join(ROOT, "tests", "golden")

# name -> (res, field, class, type, watertight_template, tets dtype, sdf dtype/shape)
CASES = {
    "sphere8_gshell": (8, "sphere", "GShell_Tets", None, True, np.int64, "f32_n1"),
    "capsule12_cloth": (12, "capsule", "hmSDF_Tets", "cloth", True, np.int64, "f32_n1"),
    "capsule12_body": (12, "capsule", "hmSDF_Tets", "body", True, np.int64, "f32_n1"),
    "adv6_gshell": (6, "adv", "GShell_Tets", None, True, np.int64, "f32_n"),
    "adv6_body": (6, "adv", "hmSDF_Tets", "body", True, np.int32, "f64_n1"),
    "adv5_open": (5, "adv", "GShell_Tets", None, False, np.int64, "f32_n"),
    "adv5_open_body": (5, "adv", "hmSDF_Tets", "body", False, np.int64, "f32_n"),
    "outside4": (4, "outside", "GShell_Tets", None, True, np.int64, "f32_n"),
    "msdfneg6": (6, "msdfneg", "GShell_Tets", None, True, np.int64, "f32_n"),
    "msdfpos6": (6, "msdfpos", "GShell_Tets", None, True, np.int64, "f32_n"),
    "smplx_layout10": (10, "smplx", "hmSDF_Tets", "cloth", True, np.int64, "f32_n1"),
    "three_faces": (0, "three", "GShell_Tets", None, True, np.int64, "f32_n"),  # torch.cross(dim=None) quirk, :19
    "single_tet": (0, "single", "GShell_Tets", None, True, np.int64, "f32_n"),
}


def make_inputs(res, field, seed=1):
    if field == "smplx":
        g = grids.smplx_layout_grid(res, dilate=0.25, seed=3)
        pos, tets = g["v"], g["f"]
        sdf, msdf = grids.capsule_garment_field(pos)
        return pos, sdf, msdf, tets
    if field in ("three", "single"):
        rng = np.random.default_rng(11)
        nt = 3 if field == "three" else 1
        pos = rng.standard_normal((4 * nt, 3)).astype(np.float32)
        tets = np.arange(4 * nt, dtype=np.int64).reshape(nt, 4)
        sdf = -np.abs(rng.standard_normal(4 * nt)).astype(np.float32) - np.float32(0.1)
        sdf[::4] = np.float32(0.7)  # one inside vertex per tet -> one triangle each
        msdf = rng.standard_normal(4 * nt).astype(np.float32)
        return pos, sdf, msdf, tets
    pos, tets = grids.kuhn_grid(res)
    if field == "sphere":
        sdf, msdf = grids.sphere_plane_field(pos)
    elif field == "capsule":
        sdf, msdf = grids.capsule_garment_field(pos)
    elif field == "adv":
        pos, sdf, msdf = grids.adversarial_field(pos, res, seed)
    elif field == "outside":
        sdf, msdf = grids.sphere_plane_field(pos)
        sdf = -np.abs(sdf) - np.float32(1)
    elif field == "msdfneg":
        sdf, msdf = grids.sphere_plane_field(pos)
        msdf = -np.abs(msdf) - np.float32(1)
    elif field == "msdfpos":
        sdf, msdf = grids.sphere_plane_field(pos)
        msdf = np.abs(msdf) + np.float32(1)
    return pos, sdf, msdf, tets


def run_case(name):
    res, field, cls, typ, wt, tdtype, sdf_kind = CASES[name]
    pos, sdf, msdf, tets = make_inputs(res, field)
    tets = tets.astype(tdtype)
    sdf_in = sdf.astype(np.float64) if sdf_kind.startswith("f64") else sdf
    if sdf_kind.endswith("n1"):
        sdf_in = sdf_in[:, None]
    tp = torch.tensor(pos, requires_grad=True)
    ts = torch.tensor(sdf_in, requires_grad=True)
    tm = torch.tensor(msdf, requires_grad=True)
    obj = load_reference_class(cls, "cpu")
    args = (tp, ts, tm, torch.from_numpy(tets)) + ((typ,) if cls == "hmSDF_Tets" else ()) + (wt,)
    verts, faces, uvs, uv_idx, v_tng, extra = obj(*args)
    assert uvs is None and uv_idx is None
    rng = np.random.default_rng(7)
    g_verts = rng.standard_normal(tuple(verts.shape)).astype(np.float32)
    g_msdf = rng.standard_normal(tuple(extra["msdf"].shape)).astype(np.float32)
    g_mwt = rng.standard_normal(tuple(extra["msdf_watertight"].shape)).astype(np.float32)
    loss = (verts * torch.tensor(g_verts)).sum() + (extra["msdf"] * torch.tensor(g_msdf)).sum() \
        + (extra["msdf_watertight"] * torch.tensor(g_mwt)).sum()
    rec = dict(pos=pos, sdf=sdf_in, msdf=msdf, tets=tets, g_verts_aug=g_verts, g_msdf=g_msdf, g_msdf_watertight=g_mwt)
    if wt:
        g_vwt = rng.standard_normal(tuple(extra["vertices_watertight"].shape)).astype(np.float32)
        loss = loss + (extra["vertices_watertight"] * torch.tensor(g_vwt)).sum()
        rec["g_vertices_watertight"] = g_vwt
    if loss.requires_grad:
        loss.backward()
    rec.update(verts_aug=verts.detach().numpy(), faces_aug=faces.numpy(), v_tng_aug=v_tng.detach().numpy())
    for k, v in extra.items():
        rec["extra_" + k] = np.asarray(v) if not torch.is_tensor(v) else v.detach().numpy()
    for nm, t in (("grad_pos", tp), ("grad_sdf", ts), ("grad_msdf", tm)):
        if t.grad is not None:  # hmSDF_Tets(type="body") detaches msdf (negated under no_grad, :261-264)
            rec[nm] = t.grad.numpy()
    rec["meta"] = np.array([cls, str(typ), str(wt)])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"{name}: Va={verts.shape[0]} Fa={faces.shape[0]} keys={len(rec)}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for n in CASES:
        run_case(n)
