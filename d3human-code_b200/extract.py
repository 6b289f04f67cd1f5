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
    seq: int = 0
    workspace: Optional[torch.Tensor] = None
    workspace_ptr: int = 0
    workspace_cap_tets: int = -1
    counts_host: Optional[torch.Tensor] = None
    counts_ptr: int = 0
    counts: Optional[_cabi.Counts] = None
    args: _cabi.ForwardArgs = field(default_factory=_cabi.ForwardArgs)
    bargs: _cabi.BackwardArgs = field(default_factory=_cabi.BackwardArgs)

    def ensure_workspace(self):
        if self.workspace_cap_tets != self.cap_tets:
            need = _cabi.lib().d3h_workspace_bytes(self.n_tets, self.n_grid, self.cap_tets)
            if self.workspace is None or self.workspace.numel() < need:
                self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
                self.workspace_ptr = self.workspace.data_ptr()
            self.workspace_cap_tets = self.cap_tets
        if self.counts_host is None:
            # pinned host memory is device-mapped (UVA): the kernel that finalises the sizes writes them here directly
            self.counts_host = torch.zeros(_cabi.COUNTS_WORDS, dtype=torch.int64).pin_memory()
            self.counts_ptr = self.counts_host.data_ptr()
            self.counts = _cabi.Counts.from_address(self.counts_ptr)


_plans: Dict[Tuple, _Plan] = {}


def _plan_for(device: torch.device, n_tets: int, n_grid: int) -> _Plan:
    key = (device.index, n_tets, n_grid)
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
    tape: torch.Tensor          # int32 slab: edges (V,2) | corners (P) | slots (P) | runs (V+1)
    tape_ptrs: Tuple[int, int, int, int]
    n_verts: int
    n_tri: int
    n_quad: int
    counts: Dict[str, int]
    launches: int
    zero_grads: Optional[Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]]

    # views used by the tests
    @property
    def tape_edges(self):
        return self.tape[:2 * self.n_verts].view(-1, 2)

    @property
    def tape_corners(self):
        p = 3 * self.n_tri + 4 * self.n_quad
        o = (self.tape_ptrs[1] - self.tape_ptrs[0]) // 4
        return self.tape[o:o + p]


def _r4(n: int) -> int:
    """round a row count up so that the next region of a slab stays 16-byte aligned"""
    return (n + 3) & ~3


_WAIT_TIMEOUT_US = 60_000_000


def forward_raw(pos: torch.Tensor, sdf: torch.Tensor, msdf: torch.Tensor, tets_i32: torch.Tensor, msdf_negate: bool,
                watertight_template: bool, tet_range: Optional[Tuple[int, int]] = None,
                want_grads: Tuple[bool, bool, bool] = (False, False, False)) -> ForwardResult:
    """One forward extraction on the current stream.  Inputs: contiguous fp32 CUDA tensors, packed int32 tets.

    The host blocks exactly once, on the sizes of the outputs (the reference blocks ~40 times per call): the kernel that
    finalises them writes d3h_counts into pinned host memory ahead of the output-writing kernels and this function
    spins on its sequence word, so the views below are built while the GPU is still finishing the call.
    If `want_grads` names any input, the dense gradient buffers of the coming backward call are allocated here and
    zero-filled by the forward call's tail (they are returned in `zero_grads`)."""
    L = _cabi.lib()
    dev = pos.device
    n_grid, n_tets = pos.shape[0], tets_i32.shape[0]
    plan = _plan_for(dev, n_tets, n_grid)
    stream = torch.cuda.current_stream(dev).cuda_stream
    launches = 0
    a = plan.args
    with torch.cuda.device(dev):
        zero_grads = None
        if any(want_grads):
            zero_grads = (torch.empty_like(pos), torch.empty_like(sdf), torch.empty_like(msdf) if want_grads[2] else None)
            a.zero_g_pos, a.zero_g_sdf = zero_grads[0].data_ptr(), zero_grads[1].data_ptr()
            a.zero_g_msdf = zero_grads[2].data_ptr() if zero_grads[2] is not None else None
        else:
            a.zero_g_pos = a.zero_g_sdf = a.zero_g_msdf = None
        a.pos, a.sdf, a.msdf, a.tets = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr(), tets_i32.data_ptr()
        a.n_grid, a.n_tets = n_grid, n_tets
        a.tet_begin, a.tet_end = (0, n_tets) if tet_range is None else tet_range
        a.msdf_negate, a.watertight_template = int(bool(msdf_negate)), int(bool(watertight_template))
        for attempt in range(6):
            plan.ensure_workspace()
            cv, cva, cfw, cfa, ct = plan.cap_v, plan.cap_va, plan.cap_fw, plan.cap_fa, plan.cap_tets
            # three slabs: float outputs, int64 faces, int32 tape
            o_vaug, o_tng, o_maug = 0, 3 * _r4(cva), 6 * _r4(cva)
            o_vwt = o_maug + _r4(cva)
            o_twt, o_mwt = o_vwt + 3 * _r4(cv), o_vwt + 6 * _r4(cv)
            fslab = torch.empty(o_mwt + _r4(cv), dtype=torch.float32, device=dev)
            islab = torch.empty(3 * (cfa + cfw), dtype=torch.int64, device=dev)
            t_corn, t_slot = 2 * _r4(cv), 2 * _r4(cv) + 4 * ct
            t_runs = t_slot + 4 * ct
            tape = torch.empty(t_runs + cv + 1, dtype=torch.int32, device=dev)
            fp, ip, tp = fslab.data_ptr(), islab.data_ptr(), tape.data_ptr()
            a.cap_valid_tets, a.cap_verts, a.cap_verts_aug, a.cap_faces_wt, a.cap_faces_aug = ct, cv, cva, cfw, cfa
            a.verts_aug, a.v_tng_aug, a.msdf_aug = fp + 4 * o_vaug, fp + 4 * o_tng, fp + 4 * o_maug
            a.verts_wt, a.v_tng_wt, a.msdf_wt = fp + 4 * o_vwt, fp + 4 * o_twt, fp + 4 * o_mwt
            a.faces_aug, a.faces_wt = ip, ip + 24 * cfa
            tape_ptrs = (tp, tp + 4 * t_corn, tp + 4 * t_slot, tp + 4 * t_runs)
            a.tape_edges, a.tape_corners, a.tape_slots, a.tape_runs = tape_ptrs
            a.workspace, a.workspace_bytes = plan.workspace_ptr, plan.workspace.numel()
            a.counts_host = plan.counts_ptr
            plan.seq += 1
            a.seq = plan.seq
            _cabi.check(L.d3h_extract_forward(C.byref(a), stream), "d3h_extract_forward")
            launches += _launches_forward(ct, zero_grads is not None)
            _cabi.check(L.d3h_wait_counts(plan.counts_ptr, plan.seq, _WAIT_TIMEOUT_US), "d3h_wait_counts")
            c = plan.counts
            fv, t1, t2, p, v, fa = c.n_valid_tets, c.n_tri_tets, c.n_quad_tets, c.n_corners, c.n_verts, c.n_faces_aug
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
        bucket_polys = tuple(c.bucket_polys)
        # next call: predict from this call's sizes (the surface moves slowly between training iterations)
        plan.cap_tets = max(_grow(fv), min(plan.cap_tets, 2 * _grow(fv)))
        plan.cap_v, plan.cap_va = _shrink(plan.cap_v, v), _shrink(plan.cap_va, va)
        plan.cap_fw, plan.cap_fa = _shrink(plan.cap_fw, fw), _shrink(plan.cap_fa, fa)
        ast = torch.as_strided
        res = ForwardResult(
            ast(fslab, (va, 3), (3, 1), o_vaug), ast(fslab, (va, 3), (3, 1), o_tng), ast(fslab, (va,), (1,), o_maug),
            ast(islab, (fa, 3), (3, 1), 0), ast(fslab, (v, 3), (3, 1), o_vwt), ast(fslab, (v, 3), (3, 1), o_twt),
            ast(fslab, (v,), (1,), o_mwt), ast(islab, (fw, 3), (3, 1), 3 * cfa), tape, tape_ptrs, v, t1, t2,
            dict(n_valid_tets=fv, n_tri_tets=t1, n_quad_tets=t2, n_corners=p, n_verts=v, n_verts_aug=va,
                 n_faces_watertight=fw, n_faces_aug=fa, bucket_polys=bucket_polys),
            launches, zero_grads)
    return res


def _shrink(cap: int, need: int) -> int:
    g = _grow(need)
    return g if (cap < g or cap > 2 * g) else cap


LAUNCHES_BACKWARD = 1  # adjoint_kernel (the zero-fill of the dense gradients rides on the forward call)


def _launches_forward(cap_tets: int, zero: bool) -> int:
    """Kernels one d3h_extract_forward enqueues: prepare, classify, compact, [bucket_scan, partition, group_sort,
    vertex_emit, poly_faces, poly_cut | publish_counts] (+ zero_block_kernel when the gradient buffers are pre-zeroed)."""
    return (4 if cap_tets <= 0 else 9) + (1 if zero else 0)


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
        need = ctx.needs_input_grad
        want = (need[0] or need[1] or need[2], need[0] or need[1] or need[2], bool(need[2]) and not msdf_negate)
        r = forward_raw(pos, sdf, msdf, tets_i32, msdf_negate, watertight_template, want_grads=want)
        ctx.save_for_backward(pos, sdf, msdf, r.tape, r.verts_wt, r.msdf_wt)
        ctx.meta = (r.n_verts, r.n_tri, r.n_quad, bool(msdf_negate), tets_i32.shape[0], r.tape_ptrs)
        ctx.zero_grads = r.zero_grads
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(r.faces_aug, r.faces_wt)
        _ExtractFn.last_counts = r.counts
        _ExtractFn.last_launches = r.launches
        return r.verts_aug, r.v_tng_aug, r.msdf_aug, r.verts_wt, r.v_tng_wt, r.msdf_wt, r.faces_aug, r.faces_wt

    @staticmethod
    def backward(ctx, g_verts_aug, g_tng_aug, g_msdf_aug, g_verts_wt, g_tng_wt, g_msdf_wt, _gfa, _gfw):
        if g_tng_aug is not None or g_tng_wt is not None:
            raise NotImplementedError(
                "gradients through v_tng (vertex tangents) are not implemented; D3-Human never uses them "
                "(hmsdf.py:454,548 drop v_tng). Detach v_tng before using it in a loss.")
        pos, sdf, msdf, tape, verts_wt, msdf_wt = ctx.saved_tensors
        n_verts, n_tri, n_quad, negate, n_tets, tape_ptrs = ctx.meta
        zg, ctx.zero_grads = ctx.zero_grads, None  # the pre-zeroed buffers serve ONE backward pass
        g_pos, g_sdf, g_msdf = backward_raw(pos, sdf, msdf, tape_ptrs, verts_wt, msdf_wt, n_verts, n_tri, n_quad,
                                            negate, n_tets, g_verts_aug, g_msdf_aug, g_verts_wt, g_msdf_wt,
                                            want_msdf=ctx.needs_input_grad[2] and not negate, prezeroed=zg)
        return g_pos, g_sdf, g_msdf, None, None, None


def backward_raw(pos, sdf, msdf, tape_ptrs, verts_wt, msdf_wt, n_verts, n_tri, n_quad, negate, n_tets,
                 g_verts_aug, g_msdf_aug, g_verts_wt, g_msdf_wt, want_msdf=True, prezeroed=None):
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
        if prezeroed is not None:
            g_pos, g_sdf, g_msdf = prezeroed
            if not want_msdf:
                g_msdf = None
            elif g_msdf is None:
                g_msdf = torch.zeros_like(msdf)
        else:
            g_pos = torch.empty_like(pos)
            g_sdf = torch.empty_like(sdf)
            g_msdf = torch.empty_like(msdf) if want_msdf else None
        b = plan.bargs
        b.pos, b.sdf, b.msdf, b.n_grid = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr(), n_grid
        b.msdf_negate = int(negate)
        b.grads_prezeroed = int(prezeroed is not None)
        b.tape_edges, b.tape_corners, b.tape_slots, b.tape_runs = tape_ptrs
        b.verts_wt, b.msdf_wt = verts_wt.data_ptr(), msdf_wt.data_ptr()
        b.n_verts, b.n_tri_tets, b.n_quad_tets = n_verts, n_tri, n_quad
        b.g_verts_aug, b.g_msdf_aug, b.g_verts_wt, b.g_msdf_wt = p_gva, p_gma, p_gvw, p_gmw
        b.g_pos, b.g_sdf = g_pos.data_ptr(), g_sdf.data_ptr()
        b.g_msdf = g_msdf.data_ptr() if g_msdf is not None else None
        b.workspace, b.workspace_bytes = None, 0
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
