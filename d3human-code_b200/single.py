"""The drop-in single call (`GShell_Tets()(...)` / `hmSDF_Tets()(...)`, one frame, the reference's calling pattern:
two calls per training iteration, hmsdf.py:454, 548) with the host work cut to what one call needs.

`extract.extract_frames*` is generic over a batch: argument blocks are numpy matrices, outputs are described by lists
of frame references, every step is vectorised over B frames -- for B = 1 that machinery costs more than the GPU work
(round 1: 0.40 ms per call of which ~0.14 ms were kernels).  This module keeps, per (device, grid, flags), ONE persistent
`d3h_forward_args` / `d3h_backward_args` pair whose static words are filled once; a call only stores the pointers that
change (inputs, two output slabs, gradient buffers, the count slot), makes one library call per direction and wraps
the outputs.  Same kernels, same plan (capacities, workspace, count ring) as the batch path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from . import _cabi
from . import extract as E

_f32 = torch.float32
_ast = torch.as_strided


def _raw_stream(dev_index: int) -> int:
    """cudaStream_t of torch's current stream on the device (the library enqueues on the caller's stream)."""
    return torch._C._cuda_getCurrentRawStream(dev_index)


if not torch.cuda.is_available():      # CPU-only test rigs (tests/emu, tests/_fake_lib) patch torch.cuda.current_stream
    def _raw_stream(dev_index: int) -> int:  # noqa: F811
        return torch.cuda.current_stream(dev_index).cuda_stream


class _State:
    """Per (device, F, N, watertight flag, negate flag) state of the single call."""
    __slots__ = ("plan", "tets", "static", "wt", "negate", "fa", "ba", "fa_ref", "ba_ref", "geom", "dev", "n_grid",
                 "n_tets", "L", "dev_index", "side")

    def __init__(self, dev, tets_i32, n_grid, static, wt, negate):
        self.dev, self.dev_index = dev, dev.index
        self.tets, self.static, self.wt, self.negate = tets_i32, static, wt, negate
        self.n_grid, self.n_tets = n_grid, tets_i32.shape[0]
        self.plan = E._plan_for(dev, self.n_tets, n_grid)
        self.L = _cabi.lib()
        fa, ba = _cabi.ForwardArgs(), _cabi.BackwardArgs()
        fa.tets = tets_i32.data_ptr()
        fa.n_grid, fa.n_tets, fa.tet_begin, fa.tet_end = n_grid, self.n_tets, 0, self.n_tets
        fa.msdf_negate, fa.watertight_template = int(negate), int(wt)
        if static is not None:
            fa.edge_off, fa.edge_ab, fa.n_edges = static[0].data_ptr(), static[1].data_ptr(), static[2]
            if static[3] is not None:
                fa.tet_edge_rank = static[3].data_ptr()
            if static[6] is not None:
                fa.etet_off, fa.etets = static[5].data_ptr(), static[6].data_ptr()
                if static[4] is not None:
                    fa.edge_b = static[4].data_ptr()
                if len(static) > 7 and static[7] is not None:
                    fa.etets8 = static[7].data_ptr()
                if len(static) > 9 and static[8] is not None:
                    fa.edge_rows, fa.edge_row_off = static[8].data_ptr(), static[9].data_ptr()
                if len(static) > 11 and static[10][0] is not None:
                    er, tr = static[10], static[11]
                    fa.edge_runs, fa.edge_run_chunk, fa.edge_run_ids = er[0].data_ptr(), er[1].data_ptr(), er[2].data_ptr()
                    fa.n_edge_runs = er[0].shape[0]
                    if tr[0] is not None:
                        fa.tet_runs, fa.tet_run_chunk, fa.tet_run_ids = tr[0].data_ptr(), tr[1].data_ptr(), tr[2].data_ptr()
                        fa.n_tet_runs = tr[0].shape[0]
        ba.n_grid, ba.msdf_negate, ba.grads_prezeroed = n_grid, int(negate), 1
        self.fa, self.ba = fa, ba
        self.fa_ref, self.ba_ref = C.addressof(fa), C.addressof(ba)
        self.geom = None
        self.side = None     # (faces_aug, faces_watertight, sizes) of the call in flight: handed over outside the node


class _Geom:
    """Slab geometry for the current capacities of the plan (byte offsets inside the two slabs of a call)."""
    __slots__ = ("caps", "ws_ptr", "ws_bytes", "f_len", "i_len", "o", "o_tape", "o_vacc", "t_corn", "t_slot", "t_runs")

    def __init__(self, st: _State):
        plan = st.plan
        cv, cva, cfw, cfa, ct = plan.cap_v, plan.cap_va, plan.cap_fw, plan.cap_fa, plan.cap_tets
        self.caps = (cv, cva, cfw, cfa, ct)
        self.ws_ptr, self.ws_bytes = plan.workspace_ptrs[0], plan.workspace_bytes
        r4 = E._r4
        o_vaug, o_tng, o_maug = 0, 3 * r4(cva), 6 * r4(cva)
        o_vwt = o_maug + r4(cva)
        o_twt, o_mwt = o_vwt + 3 * r4(cv), o_vwt + 6 * r4(cv)
        self.o = (o_vaug, o_tng, o_maug, o_vwt, o_twt, o_mwt)          # float offsets
        n = o_mwt + r4(cv)
        self.o_vacc = n                                                    # (cap_v, 8) adjoint accumulator, static paths
        if st.static is not None:
            n += 8 * r4(cv)
        # the tape (int32: edges | corners | slots | runs) lives behind the float outputs in the same slab: nothing but
        # the backward kernels ever reads it, so it needs no tensor of its own
        self.o_tape = n
        self.t_corn = 2 * r4(cv)
        self.t_slot = self.t_corn + 4 * ct
        self.t_runs = self.t_slot + 4 * ct
        n += r4(self.t_runs + cv + 1) if st.static is None else r4(self.t_slot)
        self.f_len = n
        self.i_len = 3 * (cfa + cfw)
        fa = st.fa
        fa.cap_valid_tets, fa.cap_verts, fa.cap_verts_aug, fa.cap_faces_wt, fa.cap_faces_aug = ct, cv, cva, cfw, cfa
        fa.workspace, fa.workspace_bytes = self.ws_ptr, self.ws_bytes


_states: Dict[Tuple, _State] = {}
_last = [None, None, None]      # (tet_fx4 object, its version, N) -> packed tets of the most recent call


def reset() -> None:
    _states.clear()
    _last[0] = _last[1] = _last[2] = None


def _state_for(tet_fx4, n_grid, wt, negate) -> _State:
    if _last[0] is tet_fx4 and _last[1] == tet_fx4._version and _last[2][0] == n_grid:
        tets = _last[2][1]
    else:
        tets = E.packed_tets(tet_fx4, n_grid)
        _last[0], _last[1], _last[2] = tet_fx4, tet_fx4._version, (n_grid, tets)
    static = E.static_edges_for(tets, n_grid)
    key = (tets.data_ptr(), tets._version, n_grid, wt, negate, id(static))
    st = _states.get(key)
    if st is None:
        if len(_states) > 32:
            _states.clear()
        st = _states[key] = _State(tets.device, tets, n_grid, static, wt, negate)
    return st


def _launch(st: _State, pos, sdf, msdf, need):
    """Allocate the slabs and gradient buffers of one call and enqueue its kernels.  Returns what `_finish` needs."""
    plan, fa, L = st.plan, st.fa, st.L
    static = st.static
    plan.ensure(1, static[2] if static is not None else 0)
    g = st.geom
    if (g is None or g.caps != (plan.cap_v, plan.cap_va, plan.cap_fw, plan.cap_fa, plan.cap_tets)
            or g.ws_ptr != plan.workspace_ptrs[0] or g.ws_bytes != plan.workspace_bytes):
        g = st.geom = _Geom(st)
    dev = st.dev
    fslab = torch.empty(g.f_len, dtype=_f32, device=dev)
    islab = torch.empty(g.i_len, dtype=torch.int64, device=dev)
    fb, ib = fslab.data_ptr(), islab.data_ptr()
    o = g.o
    fa.pos, fa.sdf, fa.msdf = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr()
    fa.verts_aug, fa.v_tng_aug, fa.msdf_aug = fb + 4 * o[0], fb + 4 * o[1], fb + 4 * o[2]
    fa.verts_wt, fa.v_tng_wt, fa.msdf_wt = fb + 4 * o[3], fb + 4 * o[4], fb + 4 * o[5]
    fa.faces_aug, fa.faces_wt = ib, ib + 24 * g.caps[3]
    tb = fb + 4 * g.o_tape
    fa.tape_edges, fa.tape_corners = tb, tb + 4 * g.t_corn
    if static is None:
        fa.tape_slots, fa.tape_runs = tb + 4 * g.t_slot, tb + 4 * g.t_runs
    else:
        fa.vacc = fb + 4 * g.o_vacc
    gp = gs = gm = None
    if need[0] or need[1] or need[2]:
        # dense gradient buffers of the coming backward call: allocated now, zero-filled by the tail of this forward
        # call (HBM idles behind the latency-bound surface kernels)
        gp, gs = torch.empty_like(pos), torch.empty_like(sdf)
        fa.zero_g_pos, fa.zero_g_sdf = gp.data_ptr(), gs.data_ptr()
        if need[2] and not st.negate:      # "body" never reaches msdf (hmsdf_tets_split.py:256-264)
            gm = torch.empty_like(msdf)
            fa.zero_g_msdf = gm.data_ptr()
        else:
            fa.zero_g_msdf = 0
    else:
        fa.zero_g_pos = fa.zero_g_sdf = fa.zero_g_msdf = 0
    if plan.inflight >= E._COUNT_RING:
        raise RuntimeError("the count ring of this grid is full of un-read batches: call result() on the earlier futures")
    slot = plan.slot
    plan.slot = (slot + 1) % E._COUNT_RING
    plan.seq += 1
    seq = plan.seq
    fa.counts_host = cptr = plan.counts_ptr + slot * 128
    fa.seq = seq
    if E._unjoined.get(st.dev_index):
        # the lanes of an un-joined batch share workspace 0 with this call
        stream = torch.cuda.current_stream(dev).cuda_stream
        _cabi.check(L.d3h_lanes_join(stream), "d3h_lanes_join")
        E._unjoined[st.dev_index] = False
    rs = _raw_stream(st.dev_index)
    ls = plan.last_stream
    if ls is None or ls.cuda_stream != rs:
        # workspace and count ring are per grid, not per stream: a call from another stream queues behind the last user
        cur = torch.cuda.current_stream(dev)
        if ls is not None:
            cur.wait_stream(ls)
        plan.last_stream = cur
    rc = L.d3h_extract_forward(st.fa_ref, rs)
    if rc:
        _cabi.check(rc, "d3h_extract_forward")
    return fslab, islab, g, cptr, seq, slot, (gp, gs, gm)


def _forward(st: _State, pos, sdf, msdf, need):
    """Launch, read the sizes (the one host wait of a call), regrow on overflow.  -> (slabs, geometry, sizes, grads)"""
    plan = st.plan
    L = st.L
    launches = 0
    for _attempt in range(6):
        fslab, islab, g, cptr, seq, slot, gbufs = _launch(st, pos, sdf, msdf, need)
        cv, cva, cfw, cfa, ct = g.caps
        launches += E.forward_launches(st.static, ct) + (1 if gbufs[0] is not None else 0)
        rc = L.d3h_wait_counts(cptr, seq, E._WAIT_TIMEOUT_US)
        if rc:
            _cabi.check(rc, "d3h_wait_counts")
        row = plan.counts_np[slot].tolist()
        fv, t1, t2, p, v, nfa = row[0:6]
        va, fw = v + p, t1 + 2 * t2
        if fv > ct:      # record buffer too small: the surface stages were skipped, sizes unknown; bounds that cannot overflow
            pc = 3 * t1 + 4 * t2
            plan.cap_tets = max(plan.cap_tets, E._grow(fv))
            plan.cap_v, plan.cap_va = max(plan.cap_v, pc), max(plan.cap_va, 2 * pc)
            plan.cap_fw, plan.cap_fa = max(plan.cap_fw, fw), max(plan.cap_fa, 2 * t1 + 4 * t2)
            continue
        if v > cv or va > cva or fw > cfw or nfa > cfa:
            plan.cap_v, plan.cap_va = max(plan.cap_v, E._grow(v)), max(plan.cap_va, E._grow(va))
            plan.cap_fw, plan.cap_fa = max(plan.cap_fw, E._grow(fw)), max(plan.cap_fa, E._grow(nfa))
            continue
        break
    else:  # pragma: no cover
        raise RuntimeError("d3h_extract_forward: capacities did not converge")
    # next call: predict from this call's sizes (the surface moves slowly between training iterations)
    gf = E._grow(fv)
    if not (gf <= plan.cap_tets <= 2 * gf):
        plan.cap_tets = gf
    plan.cap_v, plan.cap_va = E._shrink(plan.cap_v, v), E._shrink(plan.cap_va, va)
    plan.cap_fw, plan.cap_fa = E._shrink(plan.cap_fw, fw), E._shrink(plan.cap_fa, nfa)
    counts = (fv, t1, t2, p, v, va, fw, nfa, tuple(row[6:12]))
    return fslab, islab, g, counts, gbufs, launches


class _SingleFn(torch.autograd.Function):
    """One frame as one autograd node: forward = one library call, backward = one library call.

    Differentiable outputs: verts_aug, extra['msdf'], vertices_watertight, msdf_watertight, msdf_boundary (= msdf[V:]).
    v_tng_* are differentiable through the optional tangent branch (extract.tangent_branch; the reference's training
    never consumes them, hmsdf.py:454,548).  The int64 face arrays are handed over through `st`."""

    @staticmethod
    def forward(ctx, st, pos, sdf, msdf):
        need = ctx.needs_input_grad[1:]
        fslab, islab, g, counts, gbufs, launches = _forward(st, pos, sdf, msdf, need)
        fv, t1, t2, p, v, va, fw, nfa, buckets = counts
        o = g.o
        verts_aug = _ast(fslab, (va, 3), (3, 1), o[0])
        v_tng_aug = _ast(fslab, (va, 3), (3, 1), o[1])
        msdf_aug = _ast(fslab, (va,), (1,), o[2])
        verts_wt = _ast(fslab, (v, 3), (3, 1), o[3])
        v_tng_wt = _ast(fslab, (v, 3), (3, 1), o[4])
        msdf_wt = _ast(fslab, (v,), (1,), o[5])
        msdf_bnd = _ast(fslab, (p,), (1,), o[2] + v)
        st.side = (_ast(islab, (nfa, 3), (3, 1), 0), _ast(islab, (fw, 3), (3, 1), 3 * g.caps[3]), counts)
        ctx.st, ctx.geom, ctx.fslab, ctx.gbufs, ctx.sizes = st, g, fslab, gbufs, (v, t1, t2, va)
        ctx.islab = islab                 # faces_watertight: only read by the tangent branch of the backward pass
        ctx.save_for_backward(pos, sdf, msdf)
        ctx.set_materialize_grads(False)
        E._ExtractFn.total_launches += launches
        E._ExtractFn.last_launches = launches
        return verts_aug, v_tng_aug, msdf_aug, verts_wt, v_tng_wt, msdf_wt, msdf_bnd

    @staticmethod
    def backward(ctx, g0, g1, g2, g3, g4, g5, g6):
        st, g = ctx.st, ctx.geom
        pos, sdf, msdf = ctx.saved_tensors
        need = ctx.needs_input_grad[1:]
        gbufs, ctx.gbufs = ctx.gbufs, None       # the pre-zeroed buffers serve ONE backward pass
        if gbufs is None or gbufs[0] is None:
            if not (need[0] or need[1] or need[2]):
                raise RuntimeError("backward through an extraction whose inputs did not require gradients")
            # second backward through the same node (retain_graph): fresh zero-filled buffers
            gbufs = (torch.zeros_like(pos), torch.zeros_like(sdf),
                     torch.zeros_like(msdf) if (need[2] and not st.negate) else None)
        gp, gs, gm = gbufs
        if g0 is None and g1 is None and g2 is None and g3 is None and g4 is None and g5 is None and g6 is None:
            return None, gp if need[0] else None, gs if need[1] else None, gm if need[2] else None
        v, t1, t2, va = ctx.sizes
        ba = st.ba
        keep = []
        ba.g_verts_tng = ba.g_mvert_tng = 0
        if g1 is not None or g4 is not None:
            # optional branch: through the tangents -> extra per-vertex gradients for the adjoint kernel
            fb_ = ctx.fslab.data_ptr()
            gvt, gmt, kept = E.tangent_branch(
                st.dev, st.n_tets, fb_ + 4 * g.o[3], fb_ + 4 * g.o[5], fb_ + 4 * g.o[4], ctx.islab.data_ptr() + 24 * g.caps[3],
                fb_ + 4 * (g.o_tape + g.t_corn), v, t1, t2, g1, g4)
            keep.append((gvt, gmt, kept))
            ba.g_verts_tng, ba.g_mvert_tng = gvt.data_ptr(), gmt.data_ptr()

        def ptr(t, rows):
            if t is None:
                return 0
            if t.dtype is not _f32 or not t.is_contiguous():
                t = t.contiguous().float()
                keep.append(t)
            assert t.shape[0] == rows, (tuple(t.shape), rows)
            return t.data_ptr()

        ba.pos, ba.sdf, ba.msdf = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr()
        fb = ctx.fslab.data_ptr()
        tb = fb + 4 * g.o_tape
        ba.tape_edges, ba.tape_corners = tb, tb + 4 * g.t_corn
        if st.static is None:
            ba.tape_slots, ba.tape_runs, ba.vacc = tb + 4 * g.t_slot, tb + 4 * g.t_runs, 0
        else:
            ba.tape_slots, ba.tape_runs, ba.vacc = 0, 0, fb + 4 * g.o_vacc
        ba.verts_wt, ba.msdf_wt = fb + 4 * g.o[3], fb + 4 * g.o[5]
        ba.n_verts, ba.n_tri_tets, ba.n_quad_tets = v, t1, t2
        ba.g_verts_aug, ba.g_msdf_aug = ptr(g0, va), ptr(g2, va)
        ba.g_verts_wt, ba.g_msdf_wt, ba.g_msdf_boundary = ptr(g3, v), ptr(g5, v), ptr(g6, va - v)
        ba.g_pos, ba.g_sdf, ba.g_msdf = gp.data_ptr(), gs.data_ptr(), (gm.data_ptr() if gm is not None else 0)
        E._ExtractFn.total_launches += 2 if st.static is not None else 1
        rc = st.L.d3h_extract_backward(st.ba_ref, _raw_stream(st.dev_index))
        if rc:
            _cabi.check(rc, "d3h_extract_backward")
        return None, gp if need[0] else None, gs if need[1] else None, gm if need[2] else None


def _ok(t) -> bool:
    return t.dtype is _f32 and t.is_contiguous() and t.data_ptr() % 16 == 0


def extract(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_negate: bool = False, output_watertight_template: bool = True):
    """Shared body of GShell_Tets.__call__ / hmSDF_Tets.__call__: returns the reference's 6-tuple
    (gshell_tets.py:447, hmsdf_tets_split.py:454)."""
    if not pos_nx3.is_cuda:
        E._check_cuda(pos_nx3)
    if pos_nx3.dim() != 2 or pos_nx3.shape[1] != 3:
        raise ValueError(f"pos_nx3 must have shape (N,3), got {tuple(pos_nx3.shape)}")
    n_grid = pos_nx3.shape[0]
    pos = pos_nx3 if _ok(pos_nx3) else E._aligned(pos_nx3.float())
    # sdf_n is (N,1) from the SDF MLP or (N,); a contiguous fp32 tensor of N elements is consumed as it is (its
    # gradient comes back in its own shape), anything else goes through .float() like gshell_tets.py:254
    sdf = sdf_n if (_ok(sdf_n) and sdf_n.numel() == n_grid) else E._prep_field(sdf_n, n_grid)
    msdf = msdf_n if (_ok(msdf_n) and msdf_n.numel() == n_grid) else E._prep_field(msdf_n, n_grid)
    if not (sdf.is_cuda and msdf.is_cuda):
        E._check_cuda(sdf), E._check_cuda(msdf)
    wt = bool(output_watertight_template)
    st = _state_for(tet_fx4, n_grid, wt, bool(msdf_negate))
    verts_aug, v_tng_aug, msdf_aug, verts_wt, v_tng_wt, msdf_wt, msdf_bnd = _SingleFn.apply(st, pos, sdf, msdf)
    faces_aug, faces_wt, counts = st.side
    st.side = None
    E._ExtractFn.last_counts = counts      # expanded on demand (extract.last_counts)
    v = counts[4]
    if wt:  # gshell_tets.py:430-439
        extra = {"n_verts_watertight": v, "vertices_watertight": verts_wt, "faces_watertight": faces_wt,
                 "v_tng_watertight": v_tng_wt, "msdf": msdf_aug, "msdf_watertight": msdf_wt, "msdf_boundary": msdf_bnd}
    else:   # gshell_tets.py:440-445
        extra = {"msdf": msdf_aug, "msdf_watertight": msdf_wt, "msdf_boundary": msdf_bnd}
    return verts_aug, faces_aug, None, None, v_tng_aug, extra


def profile_scan_kernel(tet_fx4, n_grid: int, reps: int = 50, flush=None, msdf_negate: bool = False,
                        output_watertight_template: bool = True):
    """Diagnostics (bench.py `roofline`): microseconds per launch of edge_scan_kernel timed alone -- `reps` launches on
    the workspace state the LAST single call on this grid left behind, each between its own pair of CUDA events, with the
    uint8 tensor `flush` (larger than L2) filled before every launch (d3h_profile_scan_kernel).  None when the grid is not
    on the edge-scan path (yet)."""
    st = _state_for(tet_fx4, n_grid, bool(output_watertight_template), bool(msdf_negate))
    if st.static is None or st.static[6] is None or not st.fa.workspace:
        return None
    ms = C.c_float(0.0)
    rc = st.L.d3h_profile_scan_kernel(st.fa_ref, int(reps), flush.data_ptr() if flush is not None else None,
                                      flush.numel() if flush is not None else 0, C.byref(ms), _raw_stream(st.dev_index))
    if rc:
        _cabi.check(rc, "d3h_profile_scan_kernel")
    return float(ms.value) * 1e3 / reps
