"""`GShell_Tets`: drop-in for the reference class of the same name (geometry/gshell_tets.py:89-447).

    from d3human_code_b200.geometry.gshell_tets import GShell_Tets
    verts, faces, uvs, uv_idx, v_tng, extra = GShell_Tets()(v_deformed, sdf, msdf, indices)   # hmsdf.py:454-455

Constructor takes no arguments (hmsdf.py:184); the call signature, the 6-tuple, the dtypes (faces are torch.long) and
the `extra` keys are the reference's.  Slots 2 and 3 (`uvs`, `uv_idx`) are None exactly like the reference
(gshell_tets.py:447).  The look-up tables the reference uploads in __init__ (:91-190) live in the CUDA library's
constant memory; they are exposed read-only here for code that inspects them.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _cabi
from ..extract import extract

_TABLES = {  # attribute name -> (d3h_debug_table id, shape)
    "num_triangles_table": (0, (16,)),
    "num_triangles_tri_table": (4, (8,)),
    "num_triangles_quad_table": (6, (16,)),
    "triangle_table_tri": (3, (8, 6)),
    "triangle_table_quad": (5, (16, 12)),
}


def _table(which: int, shape):
    n = 1
    for s in shape:
        n *= s
    buf = (C.c_int8 * n)()
    got = _cabi.lib().d3h_debug_table(which, buf, n)
    if got != n:
        raise RuntimeError(f"d3h_debug_table({which}) returned {got}")
    return torch.tensor(list(buf), dtype=torch.long).reshape(shape)


class _TetsExtractor:
    """Shared implementation; subclasses fix the call signature."""

    def __init__(self):
        _cabi.lib()  # fail loudly at construction if the CUDA library is missing

    def __getattr__(self, name):  # lazy, read-only views of the case tables (reference: attributes set in __init__)
        if name in _TABLES:
            t = _table(*_TABLES[name])
            object.__setattr__(self, name, t)
            return t
        if name == "triangle_table":
            t = _table(2, (16, 6))
            object.__setattr__(self, name, t)
            return t
        if name == "mesh_edge_table":  # the library stores the 4 loop corners; the reference closes the loop (:110-127)
            loop = _table(1, (16, 4))
            full = torch.full((16, 6), -1, dtype=torch.long)
            for c in range(16):
                n = int((loop[c] >= 0).sum())
                if n:
                    full[c, :n] = loop[c, :n]
                    full[c, n] = loop[c, 0]
            object.__setattr__(self, name, full)
            return full
        if name == "base_tet_edges":
            p, q = _table(7, (6,)), _table(8, (6,))
            t = torch.stack([p, q], -1).reshape(-1)
            object.__setattr__(self, name, t)
            return t
        raise AttributeError(name)


class GShell_Tets(_TetsExtractor):
    def __call__(self, pos_nx3, sdf_n, msdf_n, tet_fx4, output_watertight_template=True):
        return extract(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_negate=False,
                       output_watertight_template=output_watertight_template)
