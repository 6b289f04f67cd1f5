"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/d3h_tets.h, d3h_mesh.h and d3h_mlp.h declare,
validates arguments without touching a GPU, and carries the reference's case tables."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from d3human_code_b200 import _cabi
from oracle import gshell_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("d3h_tets.h", "d3h_mesh.h", "d3h_mlp.h", "d3h_lbs.h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d3h_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()
    declared = _declared_functions()
    assert set(declared) == set(_cabi.EXPORTED_SYMBOLS), (declared, _cabi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"libd3h_tets.so does not export {name}"
    assert lib.d3h_version() == _cabi.VERSION


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct member as gcc sees include/d3h_tets.h == the ctypes mirrors in _cabi.py."""
    import subprocess
    structs = {"d3h_counts": _cabi.Counts, "d3h_forward_args": _cabi.ForwardArgs, "d3h_backward_args": _cabi.BackwardArgs,
               "d3h_mesh_counts": _cabi.MeshCounts}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "d3h_mesh.h")}"',
             'int main(void) {']
    for cname, st in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("d3h_tet_record %zu\\n", sizeof(d3h_tet_record));', '  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"
    assert int(got["d3h_tet_record"]) == _cabi.TET_RECORD_BYTES
    assert C.sizeof(_cabi.Counts) == 16 * 8      # one 128-byte slot per frame in the pinned counts buffer


def test_mesh_entry_points_validate_arguments():
    lib = _cabi.lib()
    a = lib.d3h_mesh_edges_workspace_bytes(100_000, 300_000)
    b = lib.d3h_mesh_edges_workspace_bytes(200_000, 300_000)
    assert 0 < a < b and a % 256 == 0
    assert lib.d3h_mesh_edges_workspace_bytes(-1, 10) == _cabi.D3H_E_BADARG
    assert lib.d3h_mesh_edges_workspace_bytes(1 << 30, 10) == _cabi.D3H_E_BADARG          # 3F must stay below 2^31
    assert lib.d3h_mesh_edges(None, 10, 10, None, 30, None, 0, None, None, 1, None) == _cabi.D3H_E_BADARG
    assert b"null" in lib.d3h_last_error_string()
    assert lib.d3h_mesh_wait_counts(None, 1, 10) == _cabi.D3H_E_BADARG
    assert lib.d3h_mesh_normals_forward(None, None, 10, 10, None, None, None, None) == _cabi.D3H_E_BADARG
    assert lib.d3h_mesh_normals_backward(None, None, -1, 0, None, None, None, None) == _cabi.D3H_E_BADARG
    slot = _cabi.MeshCounts(5, 0, 0, 7)                                                   # host-only: the spin wait
    assert lib.d3h_mesh_wait_counts(C.byref(slot), 7, 1000) == 0
    assert lib.d3h_mesh_wait_counts(C.byref(slot), 8, 2000) == _cabi.D3H_E_TIMEOUT


def test_workspace_bytes_contract():
    lib = _cabi.lib()
    a = lib.d3h_workspace_bytes(12_582_912, 2_146_689, 0)
    b = lib.d3h_workspace_bytes(12_582_912, 2_146_689, 100_000)
    c = lib.d3h_workspace_bytes(12_582_912, 2_146_689, 200_000)
    assert 0 < a < b < c
    assert (c - b) < 2 * (b - a)
    assert lib.d3h_workspace_bytes(-1, 10, 0) == _cabi.D3H_E_BADARG
    assert lib.d3h_backward_workspace_bytes(1000) == 0


def test_bad_arguments_are_rejected_before_any_launch():
    lib = _cabi.lib()
    assert lib.d3h_extract_forward(None, None) == _cabi.D3H_E_BADARG
    assert b"null" in lib.d3h_last_error_string()
    a = _cabi.ForwardArgs()
    a.n_grid, a.n_tets = 0, 10
    assert lib.d3h_extract_forward(C.byref(a), None) == _cabi.D3H_E_BADARG
    a.n_grid = 1 << 31
    assert lib.d3h_extract_forward(C.byref(a), None) == _cabi.D3H_E_BADARG
    a.n_grid = 100
    a.pos = a.sdf = a.msdf = a.tets = a.workspace = 4096
    a.tet_begin, a.tet_end = 0, 11  # beyond n_tets
    assert lib.d3h_extract_forward(C.byref(a), None) == _cabi.D3H_E_BADARG
    a.tet_end = 10
    a.workspace_bytes = 16
    assert lib.d3h_extract_forward(C.byref(a), None) == _cabi.D3H_E_SMALLWS
    with pytest.raises(RuntimeError, match="workspace"):
        _cabi.check(_cabi.D3H_E_SMALLWS, "d3h_extract_forward")
    assert lib.d3h_extract_backward(None, None) == _cabi.D3H_E_BADARG
    assert lib.d3h_pack_tets_i64(None, 1, 1, None, None, None) == _cabi.D3H_E_BADARG


def _table(which, shape):
    n = int(np.prod(shape))
    buf = (C.c_int8 * n)()
    assert _cabi.lib().d3h_debug_table(which, buf, n) == n
    return np.array(list(buf), dtype=np.int64).reshape(shape)


def test_case_tables_equal_oracle():
    """the build's constant tables equal the reference's element for element (the oracle's copies are
    pinned against the live reference in tests/test_oracle_vs_reference.py)."""
    assert np.array_equal(_table(0, (16,)), O.NUM_TRIANGLES_TABLE)
    assert np.array_equal(_table(1, (16, 4)), O.MESH_EDGE_TABLE[:, :4] * (np.arange(4) < np.where(O.NUM_TRIANGLES_TABLE == 2, 4, 3)[:, None])
                          + -1 * (np.arange(4) >= np.where(O.NUM_TRIANGLES_TABLE == 2, 4, 3)[:, None]))
    assert np.array_equal(_table(2, (16, 6)), O.TRIANGLE_TABLE)
    assert np.array_equal(_table(3, (8, 6)), O.TRIANGLE_TABLE_TRI)
    assert np.array_equal(_table(4, (8,)), O.NUM_TRIANGLES_TRI_TABLE)
    assert np.array_equal(_table(5, (16, 12)), O.TRIANGLE_TABLE_QUAD)
    assert np.array_equal(_table(6, (16,)), O.NUM_TRIANGLES_QUAD_TABLE)
    assert np.array_equal(np.stack([_table(7, (6,)), _table(8, (6,))], -1).reshape(-1), O.BASE_TET_EDGES)


def test_dropin_classes_expose_reference_tables():
    from d3human_code_b200.geometry.gshell_tets import GShell_Tets
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    for cls in (GShell_Tets, hmSDF_Tets):
        g = cls()
        assert np.array_equal(g.triangle_table.numpy(), O.TRIANGLE_TABLE)
        assert np.array_equal(g.mesh_edge_table.numpy(), O.MESH_EDGE_TABLE)
        assert np.array_equal(g.triangle_table_quad.numpy(), O.TRIANGLE_TABLE_QUAD)
        assert np.array_equal(g.base_tet_edges.numpy(), O.BASE_TET_EDGES)


def test_no_cpu_path():
    import torch
    from d3human_code_b200.geometry.gshell_tets import GShell_Tets
    pos = torch.zeros(8, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        GShell_Tets()(pos, torch.zeros(8), torch.zeros(8), torch.zeros(1, 4, dtype=torch.long))


def test_mlp_entry_points_validate_their_arguments():
    """include/d3h_mlp.h: shape / alignment rules are checked before anything is launched (no GPU needed)."""
    lib = _cabi.lib()
    buf = np.zeros(4096, dtype=np.float32)
    p = buf.ctypes.data + (-buf.ctypes.data % 16)
    assert lib.d3h_mlp_linear(p, 64, 0, 64, p, 256, None, 1, None, 0, p, 256, None) == 0            # M = 0: nothing to do
    assert lib.d3h_mlp_linear(p, 64, 8, 48, p, 256, None, 1, None, 0, p, 256, None) == _cabi.D3H_E_BADARG   # K % 32
    assert lib.d3h_mlp_linear(p, 64, 8, 64, p, 48, None, 1, None, 0, p, 256, None) == _cabi.D3H_E_BADARG    # N
    assert lib.d3h_mlp_linear(p + 4, 64, 8, 64, p, 256, None, 1, None, 0, p, 256, None) == _cabi.D3H_E_BADARG  # alignment
    assert lib.d3h_mlp_linear(p, 64, 8, 64, p, 256, None, 2, None, 0, p, 256, None) == _cabi.D3H_E_BADARG   # mode 2 needs y
    assert b"d3h_mlp_linear" in lib.d3h_last_error_string()
    assert lib.d3h_mlp_wgrad(p, 256, p, 64, 8, 192, 64, p, 64, None, p, 1 << 20, None) == _cabi.D3H_E_BADARG               # N not 128 / 256
    assert lib.d3h_mlp_embed(p, 8, 6, p, 32, 32, None) == _cabi.D3H_E_BADARG                                   # 39 channels do not fit
    assert lib.d3h_mlp_head(p, 256, 8, 256, p, None, 9, p, None) == _cabi.D3H_E_BADARG                         # d_out > 8
    assert lib.d3h_mlp_embed(p, 0, 6, p, 64, 64, None) == 0
    assert lib.d3h_mlp_pack_weight(p, 64, 300, 64, 0, 0, 0, 256, 64, p, None) == _cabi.D3H_E_BADARG           # n_valid > n_pad
    assert lib.d3h_mlp_packed_weight_bytes(256, 320) == 8 * 256 * 320
