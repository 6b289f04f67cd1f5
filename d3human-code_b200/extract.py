"""Host side of the extraction: capacity planning, workspace / tet-index caches and the autograd.Function.

Mirrors the contract of the reference's `GShell_Tets.__call__` (geometry/gshell_tets.py:253-447) and
`hmSDF_Tets.__call__` (geometry/hmsdf_tets_split.py:254-454): same inputs, same 6-tuple, same `extra` keys, gradients
to `pos_nx3`, `sdf_n`, `msdf_n`.  All arithmetic happens in libd3h_tets.so (hand-written sm_100a kernels); torch is used
for device memory, the current stream and autograd bookkeeping only.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import _cabi

_SLACK_NUM, _SLACK_DEN, _SLACK_ABS = 9, 8, 1024  # capacity = need * 9/8 + 1024 rows


def _grow(n: int) -> int:
    return n * _SLACK_NUM // _SLACK_DEN + _SLACK_ABS


# --------------------------------------------------------------------------------------------------
# static tet indices: converted once to packed int32x4 and range-checked (tet_fx4 is static for a whole
# training run, hmsdf.py:207-212; the reference re-reads the int64 array every call)
# --------------------------------------------------------------------------------------------------
_packed_cache: Dict[Tuple, Tuple[torch.Tensor, "weakref.ref"]] = {}


def packed_tets(tet_fx4: torch.Tensor, n_grid: int) -> torch.Tensor:
    """(F,4) integer tensor on a CUDA device -> contiguous int32 (F,4), 16-byte aligned, validated against N.
    Cached on (data_ptr, version, shape, dtype, device, N)."""
    if tet_fx4.dim() != 2 or tet_fx4.shape[1] != 4:
        raise ValueError(f"tet_fx4 must have shape (F,4), got {tuple(tet_fx4.shape)}")
    if not tet_fx4.is_cuda:
        raise RuntimeError("d3human-code_b200 has no CPU path: tet_fx4 must live on a CUDA device")
    key = (tet_fx4.data_ptr(), tet_fx4._version, tuple(tet_fx4.shape), tet_fx4.dtype, tet_fx4.device, int(n_grid))
    hit = _packed_cache.get(key)
    if hit is not None and hit[1]() is not None:
        return hit[0]
    L = _cabi.lib()
    n_tets = tet_fx4.shape[0]
    stream = torch.cuda.current_stream(tet_fx4.device).cuda_stream
    bad = torch.zeros(1, dtype=torch.int64, device=tet_fx4.device)
    with torch.cuda.device(tet_fx4.device):
        if tet_fx4.dtype == torch.int32 and tet_fx4.is_contiguous() and tet_fx4.data_ptr() % 16 == 0:
            out = tet_fx4
            _cabi.check(L.d3h_check_tets_i32(out.data_ptr(), n_tets, n_grid, bad.data_ptr(), stream), "d3h_check_tets_i32")
        else:
            src = tet_fx4.contiguous().to(torch.int64)
            out = torch.empty((n_tets, 4), dtype=torch.int32, device=tet_fx4.device)
            _cabi.check(L.d3h_pack_tets_i64(src.data_ptr(), n_tets, n_grid, out.data_ptr(), bad.data_ptr(), stream),
                        "d3h_pack_tets_i64")
    nbad = int(bad.item())  # one-time sync per grid
    if nbad:
        raise IndexError(f"tet_fx4 holds {nbad} vertex indices outside [0, {n_grid})")
    if len(_packed_cache) > 16:
        for k in [k for k, v in _packed_cache.items() if v[1]() is None]:
            del _packed_cache[k]
    try:
        ref = weakref.ref(tet_fx4)
    except TypeError:  # pragma: no cover
        ref = lambda: tet_fx4  # noqa: E731
    _packed_cache[key] = (out, ref)
    return out


# --------------------------------------------------------------------------------------------------
# per-(device, F, N) plan: workspace + capacities predicted from the previous call
# --------------------------------------------------------------------------------------------------
@dataclass
class _Plan:
    device: torch.device
    n_tets: int
    n_grid: int
    cap_tets: int = 0
    cap_v: int = 0
    cap_va: int = 0
    cap_fw: int = 0
    cap_fa: int = 0
    workspace: Optional[torch.Tensor] = None
    counts_host: Optional[torch.Tensor] = None
    bwd_workspace: Optional[torch.Tensor] = None
    args: _cabi.ForwardArgs = field(default_factory=_cabi.ForwardArgs)

    def ensure_workspace(self):
        need = _cabi.lib().d3h_workspace_bytes(self.n_tets, self.n_grid, self.cap_tets)
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        if self.counts_host is None:
            self.counts_host = torch.zeros(_cabi.COUNTS_WORDS, dtype=torch.int64).pin_memory()

    def ensure_bwd_workspace(self, n_verts: int) -> torch.Tensor:
        need = _cabi.lib().d3h_backward_workspace_bytes(n_verts)
        if self.bwd_workspace is None or self.bwd_workspace.numel() < need:
            self.bwd_workspace = torch.empty(_grow(need), dtype=torch.uint8, device=self.device)
        return self.bwd_workspace


_plans: Dict[Tuple, _Plan] = {}


def _plan_for(device: torch.device, n_tets: int, n_grid: int) -> _Plan:
    key = (device.type, device.index, n_tets, n_grid)
    p = _plans.get(key)
    if p is None:
        p = _plans[key] = _Plan(device=device, n_tets=n_tets, n_grid=n_grid)
    return p


def reset_plans() -> None:
    """Drop cached workspaces / capacity predictions (tests)."""
    _plans.clear()
    _packed_cache.clear()


@dataclass
class ForwardResult:
    verts_aug: torch.Tensor
    v_tng_aug: torch.Tensor
    msdf_aug: torch.Tensor
    faces_aug: torch.Tensor
    verts_wt: torch.Tensor
    v_tng_wt: torch.Tensor
    msdf_wt: torch.Tensor
    faces_wt: torch.Tensor
    tape_edges: torch.Tensor
    tape_corners: torch.Tensor
    n_verts: int
    n_tri: int
    n_quad: int
    counts: Dict[str, int]
    launches: int


def forward_raw(pos: torch.Tensor, sdf: torch.Tensor, msdf: torch.Tensor, tets_i32: torch.Tensor, msdf_negate: bool,
                watertight_template: bool, tet_range: Optional[Tuple[int, int]] = None) -> ForwardResult:
    """One forward extraction on the current stream.  Inputs: contiguous fp32 CUDA tensors, packed int32 tets.
    Synchronises the stream once to learn the output sizes (the reference syncs ~40 times per call)."""
    L = _cabi.lib()
    dev = pos.device
    n_grid, n_tets = pos.shape[0], tets_i32.shape[0]
    plan = _plan_for(dev, n_tets, n_grid)
    stream = torch.cuda.current_stream(dev)
    launches = 0
    with torch.cuda.device(dev):
        for attempt in range(6):
            plan.ensure_workspace()
            cv, cva, cfw, cfa, ct = plan.cap_v, plan.cap_va, plan.cap_fw, plan.cap_fa, plan.cap_tets
            f32 = dict(dtype=torch.float32, device=dev)
            verts_aug = torch.empty((cva, 3), **f32)
            v_tng_aug = torch.empty((cva, 3), **f32)
            msdf_aug = torch.empty((cva,), **f32)
            faces_aug = torch.empty((cfa, 3), dtype=torch.int64, device=dev)
            verts_wt = torch.empty((cv, 3), **f32)
            v_tng_wt = torch.empty((cv, 3), **f32)
            msdf_wt = torch.empty((cv,), **f32)
            faces_wt = torch.empty((cfw, 3), dtype=torch.int64, device=dev)
            tape_edges = torch.empty((cv, 2), dtype=torch.int32, device=dev)
            tape_corners = torch.empty((max(4 * ct, 1),), dtype=torch.int32, device=dev)
            a = plan.args
            a.pos, a.sdf, a.msdf, a.tets = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr(), tets_i32.data_ptr()
            a.n_grid, a.n_tets = n_grid, n_tets
            a.tet_begin, a.tet_end = (0, n_tets) if tet_range is None else tet_range
            a.msdf_negate, a.watertight_template = int(bool(msdf_negate)), int(bool(watertight_template))
            a.cap_valid_tets, a.cap_verts, a.cap_verts_aug, a.cap_faces_wt, a.cap_faces_aug = ct, cv, cva, cfw, cfa
            a.verts_aug, a.v_tng_aug, a.msdf_aug = verts_aug.data_ptr(), v_tng_aug.data_ptr(), msdf_aug.data_ptr()
            a.faces_aug, a.verts_wt, a.v_tng_wt = faces_aug.data_ptr(), verts_wt.data_ptr(), v_tng_wt.data_ptr()
            a.msdf_wt, a.faces_wt = msdf_wt.data_ptr(), faces_wt.data_ptr()
            a.tape_edges, a.tape_corners = tape_edges.data_ptr(), tape_corners.data_ptr()
            a.workspace, a.workspace_bytes = plan.workspace.data_ptr(), plan.workspace.numel()
            a.counts_host = plan.counts_host.data_ptr()
            _cabi.check(L.d3h_extract_forward(C.byref(a), stream.cuda_stream), "d3h_extract_forward")
            launches += _launches_forward(n_grid, ct)
            stream.synchronize()
            c = plan.counts_host.tolist()
            fv, t1, t2, p, v, fa = c[0], c[1], c[2], c[3], c[4], c[5]
            if fv > ct:  # record buffer too small: surface stages were skipped, sizes below are not known yet
                plan.cap_tets = _grow(fv)
                # upper bounds that cannot overflow, so the next attempt is final
                plan.cap_v, plan.cap_va = max(cv, p), max(cva, 2 * p)
                plan.cap_fw, plan.cap_fa = max(cfw, t1 + 2 * t2), max(cfa, 2 * t1 + 4 * t2)
                continue
            va, fw = v + p, t1 + 2 * t2
            if v > cv or va > cva or fw > cfw or fa > cfa:
                plan.cap_v, plan.cap_va = max(cv, _grow(v)), max(cva, _grow(va))
                plan.cap_fw, plan.cap_fa = max(cfw, _grow(fw)), max(cfa, _grow(fa))
                continue
            break
        else:  # pragma: no cover
            raise RuntimeError("d3h_extract_forward: capacities did not converge")
        # next call: predict from this call's sizes (the surface moves slowly between training iterations)
        plan.cap_tets = max(_grow(fv), min(plan.cap_tets, 2 * _grow(fv)))
        plan.cap_v, plan.cap_va = _shrink(plan.cap_v, v), _shrink(plan.cap_va, va)
        plan.cap_fw, plan.cap_fa = _shrink(plan.cap_fw, fw), _shrink(plan.cap_fa, fa)
    counts = dict(n_valid_tets=fv, n_tri_tets=t1, n_quad_tets=t2, n_corners=p, n_verts=v, n_verts_aug=va,
                  n_faces_watertight=fw, n_faces_aug=fa, bucket_polys=tuple(c[6:12]))
    return ForwardResult(verts_aug[:va], v_tng_aug[:va], msdf_aug[:va], faces_aug[:fa], verts_wt[:v], v_tng_wt[:v],
                         msdf_wt[:v], faces_wt[:fw], tape_edges[:v], tape_corners[:max(p, 0)], v, t1, t2, counts,
                         launches)


def _shrink(cap: int, need: int) -> int:
    g = _grow(need)
    return g if (cap < g or cap > 2 * g) else cap


def _launches_forward(n_grid: int, cap_tets: int) -> int:
    """Kernels one d3h_extract_forward enqueues: prepare, classify, compact, [partition, local_sort, rle_interp,
    poly_faces, vertex_frame,] poly_cut."""
    return 4 if cap_tets <= 0 else 9


# --------------------------------------------------------------------------------------------------
# autograd
# --------------------------------------------------------------------------------------------------
class _ExtractFn(torch.autograd.Function):
    """forward: (pos, sdf, msdf) -> 7 float outputs + 2 index outputs; backward: dense grads for pos / sdf / msdf.

    Differentiable: verts_aug, msdf (augmented, stop-grad coefficients), vertices_watertight, msdf_watertight.
    v_tng_* are returned for API parity but are not differentiated (the reference's own training never consumes
    them, hmsdf.py:454,548); asking for their gradient raises instead of silently returning zeros.
    """

    @staticmethod
    def forward(ctx, pos, sdf, msdf, tets_i32, msdf_negate, watertight_template):
        r = forward_raw(pos, sdf, msdf, tets_i32, msdf_negate, watertight_template)
        ctx.save_for_backward(pos, sdf, msdf, r.tape_edges, r.tape_corners, r.verts_wt, r.msdf_wt)
        ctx.meta = (r.n_verts, r.n_tri, r.n_quad, bool(msdf_negate), tets_i32.shape[0])
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(r.faces_aug, r.faces_wt)
        ctx.counts = r.counts
        _ExtractFn.last_counts = r.counts
        _ExtractFn.last_launches = r.launches
        return r.verts_aug, r.v_tng_aug, r.msdf_aug, r.verts_wt, r.v_tng_wt, r.msdf_wt, r.faces_aug, r.faces_wt

    @staticmethod
    def backward(ctx, g_verts_aug, g_tng_aug, g_msdf_aug, g_verts_wt, g_tng_wt, g_msdf_wt, _gfa, _gfw):
        if g_tng_aug is not None or g_tng_wt is not None:
            raise NotImplementedError(
                "gradients through v_tng (vertex tangents) are not implemented; D3-Human never uses them "
                "(hmsdf.py:454,548 drop v_tng). Detach v_tng before using it in a loss.")
        pos, sdf, msdf, tape_edges, tape_corners, verts_wt, msdf_wt = ctx.saved_tensors
        n_verts, n_tri, n_quad, negate, n_tets = ctx.meta
        g_pos, g_sdf, g_msdf = backward_raw(pos, sdf, msdf, tape_edges, tape_corners, verts_wt, msdf_wt, n_verts, n_tri,
                                            n_quad, negate, n_tets, g_verts_aug, g_msdf_aug, g_verts_wt, g_msdf_wt,
                                            want_msdf=ctx.needs_input_grad[2] and not negate)
        return g_pos, g_sdf, g_msdf, None, None, None


def backward_raw(pos, sdf, msdf, tape_edges, tape_corners, verts_wt, msdf_wt, n_verts, n_tri, n_quad, negate, n_tets,
                 g_verts_aug, g_msdf_aug, g_verts_wt, g_msdf_wt, want_msdf=True):
    L = _cabi.lib()
    dev = pos.device
    n_grid = pos.shape[0]
    plan = _plan_for(dev, n_tets, n_grid)

    def ptr(t, shape):
        if t is None:
            return None, None
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.contiguous().float()
        assert tuple(t.shape) == shape, (tuple(t.shape), shape)
        return t, t.data_ptr()

    va = n_verts + 3 * n_tri + 4 * n_quad
    with torch.cuda.device(dev):
        g_verts_aug, p_gva = ptr(g_verts_aug, (va, 3))
        g_msdf_aug, p_gma = ptr(g_msdf_aug, (va,))
        g_verts_wt, p_gvw = ptr(g_verts_wt, (n_verts, 3))
        g_msdf_wt, p_gmw = ptr(g_msdf_wt, (n_verts,))
        g_pos = torch.empty_like(pos)
        g_sdf = torch.empty_like(sdf)
        g_msdf = torch.empty_like(msdf) if want_msdf else None
        ws = plan.ensure_bwd_workspace(n_verts)
        b = _cabi.BackwardArgs()
        b.pos, b.sdf, b.msdf, b.n_grid = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr(), n_grid
        b.msdf_negate = int(negate)
        b.tape_edges, b.tape_corners = tape_edges.data_ptr(), tape_corners.data_ptr()
        b.verts_wt, b.msdf_wt = verts_wt.data_ptr(), msdf_wt.data_ptr()
        b.n_verts, b.n_tri_tets, b.n_quad_tets = n_verts, n_tri, n_quad
        b.g_verts_aug, b.g_msdf_aug, b.g_verts_wt, b.g_msdf_wt = p_gva, p_gma, p_gvw, p_gmw
        b.g_pos, b.g_sdf = g_pos.data_ptr(), g_sdf.data_ptr()
        b.g_msdf = g_msdf.data_ptr() if g_msdf is not None else None
        b.workspace, b.workspace_bytes = ws.data_ptr(), ws.numel()
        _cabi.check(L.d3h_extract_backward(C.byref(b), torch.cuda.current_stream(dev).cuda_stream),
                    "d3h_extract_backward")
    return g_pos, g_sdf, g_msdf


_ExtractFn.last_counts = None
_ExtractFn.last_launches = 0


def last_counts() -> Optional[Dict[str, int]]:
    """Sizes of the most recent extraction (Fv, T1, T2, P, V, Va, Fw, Fa, bucket sizes)."""
    return _ExtractFn.last_counts


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (the kernels use 16-byte vector loads); a view at an odd offset is cloned."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


def extract(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_negate: bool = False, output_watertight_template: bool = True):
    """Shared body of GShell_Tets.__call__ / hmSDF_Tets.__call__: returns the reference's 6-tuple."""
    if not pos_nx3.is_cuda:
        raise RuntimeError("d3human-code_b200 has no CPU path: inputs must live on a CUDA device "
                           "(the reference hard-codes device='cuda' as well, gshell_tets.py:108)")
    n_grid = pos_nx3.shape[0]
    if pos_nx3.dim() != 2 or pos_nx3.shape[1] != 3:
        raise ValueError(f"pos_nx3 must have shape (N,3), got {tuple(pos_nx3.shape)}")
    sdf = sdf_n.float().reshape(-1)       # gshell_tets.py:254 (.float()); (N,1) from the SDF MLP or (N,)
    msdf = msdf_n.float().reshape(-1)
    if sdf.shape[0] != n_grid or msdf.shape[0] != n_grid:
        raise ValueError("sdf_n / msdf_n must have one value per grid vertex")
    pos, sdf, msdf = _aligned(pos_nx3.float()), _aligned(sdf), _aligned(msdf)
    tets = packed_tets(tet_fx4, n_grid)
    verts_aug, v_tng_aug, msdf_aug, verts_wt, v_tng_wt, msdf_wt, faces_aug, faces_wt = _ExtractFn.apply(
        pos, sdf, msdf, tets, bool(msdf_negate), bool(output_watertight_template))
    n_wt = verts_wt.shape[0]
    if output_watertight_template:  # gshell_tets.py:430-439
        extra = {
            "n_verts_watertight": n_wt,
            "vertices_watertight": verts_wt,
            "faces_watertight": faces_wt,
            "v_tng_watertight": v_tng_wt,
            "msdf": msdf_aug,
            "msdf_watertight": msdf_wt,
            "msdf_boundary": msdf_aug[n_wt:],
        }
    else:  # gshell_tets.py:440-445
        extra = {"msdf": msdf_aug, "msdf_watertight": msdf_wt, "msdf_boundary": msdf_aug[n_wt:]}
    return verts_aug, faces_aug, None, None, v_tng_aug, extra
