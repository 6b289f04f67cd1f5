"""Generate tests/golden/mesh_*.npz from the LIVE reference `render/mesh.py` (run in the build container only).

    python oracle/make_golden_mesh.py

Each fixture holds a small triangle mesh (pos, faces), the reference's `Mesh(...).edges` (render/mesh.py:240-250),
`auto_normals(...).v_nrm` (:418-446) and the autograd gradient of sum(v_nrm * g_nrm) w.r.t. pos for a seeded g_nrm.
They pin oracle/mesh_oracle.py (tests/test_mesh_oracle.py) and the CUDA kernels (tests/test_zz_mesh.py) on machines
where the reference tree does not exist.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference_mesh  # noqa: E402

warnings.filterwarnings("ignore")
OUT = os.path.join(ROOT, "tests", "golden")


def case_inputs(name):
    """-> pos (V,3) f32, faces (F,3) i64.  Also imported by the tests to rebuild the same inputs."""
    rng = np.random.default_rng(sum(map(ord, name)))
    if name == "mesh_extracted":  # an extracted open surface: most augmented vertices are unused (zero rows, no faces)
        z = np.load(os.path.join(OUT, "capsule12_cloth.npz"))
        return z["verts_aug"].astype(np.float32), z["faces_aug"].astype(np.int64)
    if name == "mesh_watertight":
        z = np.load(os.path.join(OUT, "sphere8_gshell.npz"))
        return z["extra_vertices_watertight"].astype(np.float32), z["extra_faces_watertight"].astype(np.int64)
    if name == "mesh_soup":  # random triangles: repeated edges, repeated faces, degenerate faces (a == b), isolated vertices
        pos = rng.standard_normal((60, 3)).astype(np.float32)
        faces = rng.integers(0, 50, size=(200, 3)).astype(np.int64)
        faces[10] = faces[11]
        faces[20, 1] = faces[20, 0]
        faces[21] = [7, 7, 7]
        return pos, faces
    if name == "mesh_three":  # exactly three faces: torch.cross without dim crosses along the face axis
        pos = rng.standard_normal((5, 3)).astype(np.float32)
        return pos, np.array([[0, 1, 2], [2, 1, 3], [3, 1, 4]], np.int64)
    if name == "mesh_fan":  # one hub with 120 spokes (the hub is the smallest index: one long neighbour segment)
        n = 120
        ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
        pos = np.concatenate([[[0, 0, 0.3]], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)], 0).astype(np.float32)
        ring = 1 + np.arange(n)
        faces = np.stack([np.zeros(n, np.int64), ring, 1 + (np.arange(n) + 1) % n], 1).astype(np.int64)
        return pos, faces[rng.permutation(n)]
    if name == "mesh_single":
        return rng.standard_normal((3, 3)).astype(np.float32), np.array([[2, 0, 1]], np.int64)
    if name == "mesh_empty":
        return rng.standard_normal((4, 3)).astype(np.float32), np.zeros((0, 3), np.int64)
    if name == "mesh_flat":  # coplanar duplicate faces with opposite winding: sums cancel exactly -> (0,0,1)
        pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5]], np.float32)
        return pos, np.array([[0, 1, 2], [0, 2, 1], [1, 3, 2], [1, 2, 3]], np.int64)
    raise KeyError(name)


CASES = ("mesh_extracted", "mesh_watertight", "mesh_soup", "mesh_three", "mesh_fan", "mesh_single", "mesh_empty",
         "mesh_flat")


def upstream(name, n_verts):
    return np.random.default_rng(1000 + sum(map(ord, name))).standard_normal((n_verts, 3)).astype(np.float32)


def main():
    ref = load_reference_mesh("cpu")
    for name in CASES:
        pos, faces = case_inputs(name)
        p = torch.tensor(pos, requires_grad=True)
        f = torch.tensor(faces)
        out = {"pos": pos, "faces": faces}
        if faces.shape[0]:
            m = ref.Mesh(p, f)
            out["edges"] = m.edges.numpy()
            nm = ref.auto_normals(m)
            g = upstream(name, pos.shape[0])
            (nm.v_nrm * torch.tensor(g)).sum().backward()
            out["v_nrm"] = nm.v_nrm.detach().numpy()
            out["g_nrm"] = g
            out["g_pos"] = p.grad.numpy()
        else:  # the reference cannot build an empty mesh through torch.cat of empty index tensors on every version
            out["edges"] = np.zeros((0, 2), np.int64)
            out["v_nrm"] = np.tile(np.array([0, 0, 1], np.float32), (pos.shape[0], 1))
            out["g_nrm"] = upstream(name, pos.shape[0])
            out["g_pos"] = np.zeros_like(pos)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "V", pos.shape[0], "F", faces.shape[0], "E", out["edges"].shape[0])


if __name__ == "__main__":
    main()
