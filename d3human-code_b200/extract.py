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
from typing import Any, Dict, List, Optional, Tuple

import torch

from . import _cabi

ctypes_array = Any

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
# per-(device, F, N) plan: workspaces (one per lane) + capacities predicted from the previous call
# --------------------------------------------------------------------------------------------------
MAX_LANES = 8
DEFAULT_LANES = 4


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
    workspaces: List[torch.Tensor] = field(default_factory=list)   # one per lane
    workspace_bytes: int = 0
    workspace_cap_tets: int = -1
    counts_host: Optional[torch.Tensor] = None                     # pinned, one 128-byte slot per frame of a batch
    counts: List[_cabi.Counts] = field(default_factory=list)
    fargs: Dict[int, ctypes_array] = field(default_factory=dict)   # batch size -> (ForwardArgs * B)()
    bargs: Dict[int, ctypes_array] = field(default_factory=dict)

    def ensure(self, n_frames: int, lanes: int):
        if self.workspace_cap_tets != self.cap_tets:
            need = _cabi.lib().d3h_workspace_bytes(self.n_tets, self.n_grid, self.cap_tets)
            if need > self.workspace_bytes:
                self.workspaces = []
                self.workspace_bytes = need
            self.workspace_cap_tets = self.cap_tets
        while len(self.workspaces) < lanes:
            self.workspaces.append(torch.empty(self.workspace_bytes, dtype=torch.uint8, device=self.device))
        if len(self.counts) < n_frames:
            # pinned host memory is device-mapped (UVA): the kernel that finalises the sizes writes them here directly
            slots = max(n_frames, 2 * len(self.counts), 4)
            self.counts_host = torch.zeros(slots * _cabi.COUNTS_WORDS, dtype=torch.int64).pin_memory()
            base = self.counts_host.data_ptr()
            self.counts = [_cabi.Counts.from_address(base + i * C.sizeof(_cabi.Counts)) for i in range(slots)]
        if n_frames not in self.fargs:
            self.fargs[n_frames] = (_cabi.ForwardArgs * n_frames)()
            self.bargs[n_frames] = (_cabi.BackwardArgs * n_frames)()


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
    tape: torch.Tensor          # int32 slab of the batch; this frame: edges (V,2) | corners (P) | slots (P) | runs (V+1)
    fslab: torch.Tensor         # float slab of the batch (every float output is a view of it)
    tape_ptrs: Tuple[int, int, int, int]
    n_verts: int
    n_tri: int
    n_quad: int
    counts: Dict[str, int]

    # views used by the tests
    @property
    def tape_edges(self):
        o = (self.tape_ptrs[0] - self.tape.data_ptr()) // 4
        return self.tape[o:o + 2 * self.n_verts].view(-1, 2)

    @property
    def tape_corners(self):
        p = 3 * self.n_tri + 4 * self.n_quad
        o = (self.tape_ptrs[1] - self.tape.data_ptr()) // 4
        return self.tape[o:o + p]


def _r4(n: int) -> int:
    """round a row count up so that the next region of a slab stays 16-byte aligned"""
    return (n + 3) & ~3


_WAIT_TIMEOUT_US = 60_000_000

#: kernels one forward extraction enqueues: prepare, classify, compact, bucket_scan, partition, group_sort, vertex_emit,
#: poly_faces, poly_cut (+ zero_block_kernel when the gradient buffers are pre-zeroed); backward: adjoint_kernel
LAUNCHES_FORWARD = 9
LAUNCHES_BACKWARD = 1


def forward_frames_raw(frames, tets_i32: torch.Tensor, watertight_template: bool, lanes: int = DEFAULT_LANES,
                       zero: Optional[List[Tuple[Optional[int], Optional[int], Optional[int]]]] = None):
    """Forward extraction of a batch of frames in ONE library call (d3h_extract_forward_batch).

    frames: [(pos, sdf, msdf, msdf_negate)] -- contiguous, 16-byte aligned fp32 CUDA tensors; tensors may be shared
    between frames.  zero: per frame, the device pointers of dense gradient buffers (pos, sdf, msdf) that the call
    should zero-fill for the coming backward pass (None = nothing).
    Frames run on `lanes` concurrent lanes inside the library; the host blocks once per frame on that frame's sizes
    (written to pinned host memory by the kernel that finalises them) -- the reference blocks ~40 times per frame.
    Returns ([ForwardResult], launches)."""
    L = _cabi.lib()
    B = len(frames)
    pos0 = frames[0][0]
    dev = pos0.device
    n_grid, n_tets = pos0.shape[0], tets_i32.shape[0]
    plan = _plan_for(dev, n_tets, n_grid)
    lanes = max(1, min(int(lanes), MAX_LANES, B))
    stream = torch.cuda.current_stream(dev).cuda_stream
    launches = 0
    tets_ptr = tets_i32.data_ptr()
    wt = int(bool(watertight_template))
    with torch.cuda.device(dev):
        for attempt in range(6):
            plan.ensure(B, lanes)
            fa = plan.fargs[B]
            cv, cva, cfw, cfa, ct = plan.cap_v, plan.cap_va, plan.cap_fw, plan.cap_fa, plan.cap_tets
            # three slabs for the whole batch: float outputs, int64 faces, int32 tape; frame i owns slice i of each
            o_vaug, o_tng, o_maug = 0, 3 * _r4(cva), 6 * _r4(cva)
            o_vwt = o_maug + _r4(cva)
            o_twt, o_mwt = o_vwt + 3 * _r4(cv), o_vwt + 6 * _r4(cv)
            f_len = o_mwt + _r4(cv)
            i_len = 3 * (cfa + cfw) + (3 * (cfa + cfw)) % 2      # keep every frame's int64 slice 16-byte aligned
            t_corn, t_slot = 2 * _r4(cv), 2 * _r4(cv) + 4 * ct
            t_runs = t_slot + 4 * ct
            t_len = _r4(t_runs + cv + 1)
            fslab = torch.empty(B * f_len, dtype=torch.float32, device=dev)
            islab = torch.empty(B * i_len, dtype=torch.int64, device=dev)
            tape = torch.empty(B * t_len, dtype=torch.int32, device=dev)
            fp0, ip0, tp0 = fslab.data_ptr(), islab.data_ptr(), tape.data_ptr()
            ws_bytes = plan.workspace_bytes
            counts_base = plan.counts_host.data_ptr()
            seq0 = plan.seq
            plan.seq += B
            tape_ptrs = []
            for i, (pos, sdf, msdf, negate) in enumerate(frames):
                a = fa[i]
                fp, ip, tp = fp0 + 4 * i * f_len, ip0 + 8 * i * i_len, tp0 + 4 * i * t_len
                a.pos, a.sdf, a.msdf, a.tets = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr(), tets_ptr
                a.n_grid, a.n_tets, a.tet_begin, a.tet_end = n_grid, n_tets, 0, n_tets
                a.msdf_negate, a.watertight_template = int(bool(negate)), wt
                a.cap_valid_tets, a.cap_verts, a.cap_verts_aug, a.cap_faces_wt, a.cap_faces_aug = ct, cv, cva, cfw, cfa
                a.verts_aug, a.v_tng_aug, a.msdf_aug = fp + 4 * o_vaug, fp + 4 * o_tng, fp + 4 * o_maug
                a.verts_wt, a.v_tng_wt, a.msdf_wt = fp + 4 * o_vwt, fp + 4 * o_twt, fp + 4 * o_mwt
                a.faces_aug, a.faces_wt = ip, ip + 24 * cfa
                tps = (tp, tp + 4 * t_corn, tp + 4 * t_slot, tp + 4 * t_runs)
                tape_ptrs.append(tps)
                a.tape_edges, a.tape_corners, a.tape_slots, a.tape_runs = tps
                z = zero[i] if zero is not None else (None, None, None)
                a.zero_g_pos, a.zero_g_sdf, a.zero_g_msdf = z
                a.workspace, a.workspace_bytes = plan.workspaces[i % lanes].data_ptr(), ws_bytes
                a.counts_host = counts_base + i * 128
                a.seq = seq0 + 1 + i
                launches += (4 if ct <= 0 else LAUNCHES_FORWARD) + (1 if any(p is not None for p in z) else 0)
            _cabi.check(L.d3h_extract_forward_batch(fa, B, lanes, stream), "d3h_extract_forward_batch")
            sizes = []
            grow_tets = grow_out = False
            for i in range(B):
                _cabi.check(L.d3h_wait_counts(counts_base + i * 128, seq0 + 1 + i, _WAIT_TIMEOUT_US), "d3h_wait_counts")
                c = plan.counts[i]
                fv, t1, t2, p, v, nfa = c.n_valid_tets, c.n_tri_tets, c.n_quad_tets, c.n_corners, c.n_verts, c.n_faces_aug
                sizes.append((fv, t1, t2, p, v, nfa, tuple(c.bucket_polys)))
                if fv > ct:
                    grow_tets = True
                elif v > cv or v + p > cva or t1 + 2 * t2 > cfw or nfa > cfa:
                    grow_out = True
            if grow_tets:  # record buffer too small: surface stages were skipped for some frame, its sizes are unknown
                fv = max(s[0] for s in sizes)
                t1, t2 = max(s[1] for s in sizes), max(s[2] for s in sizes)
                p = 3 * t1 + 4 * t2
                plan.cap_tets = _grow(fv)
                # upper bounds that cannot overflow, so the next attempt is final
                plan.cap_v, plan.cap_va = max(cv, p), max(cva, 2 * p)
                plan.cap_fw, plan.cap_fa = max(cfw, t1 + 2 * t2), max(cfa, 2 * t1 + 4 * t2)
                continue
            if grow_out:
                v, va = max(s[4] for s in sizes), max(s[4] + s[3] for s in sizes)
                fw, nfa = max(s[1] + 2 * s[2] for s in sizes), max(s[5] for s in sizes)
                plan.cap_v, plan.cap_va = max(cv, _grow(v)), max(cva, _grow(va))
                plan.cap_fw, plan.cap_fa = max(cfw, _grow(fw)), max(cfa, _grow(nfa))
                continue
            break
        else:  # pragma: no cover
            raise RuntimeError("d3h_extract_forward_batch: capacities did not converge")
        # next call: predict from this call's sizes (the surface moves slowly between training iterations)
        fv = max(s[0] for s in sizes)
        v, va = max(s[4] for s in sizes), max(s[4] + s[3] for s in sizes)
        fw, nfa = max(s[1] + 2 * s[2] for s in sizes), max(s[5] for s in sizes)
        plan.cap_tets = max(_grow(fv), min(plan.cap_tets, 2 * _grow(fv)))
        plan.cap_v, plan.cap_va = _shrink(plan.cap_v, v), _shrink(plan.cap_va, va)
        plan.cap_fw, plan.cap_fa = _shrink(plan.cap_fw, fw), _shrink(plan.cap_fa, nfa)
        ast = torch.as_strided
        results = []
        for i, (fv, t1, t2, p, v, nfa, buckets) in enumerate(sizes):
            va, fw = v + p, t1 + 2 * t2
            fo, io = i * f_len, i * i_len
            results.append(ForwardResult(
                ast(fslab, (va, 3), (3, 1), fo + o_vaug), ast(fslab, (va, 3), (3, 1), fo + o_tng),
                ast(fslab, (va,), (1,), fo + o_maug), ast(islab, (nfa, 3), (3, 1), io),
                ast(fslab, (v, 3), (3, 1), fo + o_vwt), ast(fslab, (v, 3), (3, 1), fo + o_twt),
                ast(fslab, (v,), (1,), fo + o_mwt), ast(islab, (fw, 3), (3, 1), io + 3 * cfa), tape, fslab,
                tape_ptrs[i], v, t1, t2,
                dict(n_valid_tets=fv, n_tri_tets=t1, n_quad_tets=t2, n_corners=p, n_verts=v, n_verts_aug=va,
                     n_faces_watertight=fw, n_faces_aug=nfa, bucket_polys=buckets)))
    return results, launches


def forward_raw(pos, sdf, msdf, tets_i32, msdf_negate, watertight_template) -> ForwardResult:
    """One forward extraction without autograd (tests, profiling scripts)."""
    res, _ = forward_frames_raw([(pos, sdf, msdf, msdf_negate)], tets_i32, watertight_template, lanes=1)
    return res[0]


def _shrink(cap: int, need: int) -> int:
    g = _grow(need)
    return g if (cap < g or cap > 2 * g) else cap


# --------------------------------------------------------------------------------------------------
# autograd
# --------------------------------------------------------------------------------------------------
_OUTS_PER_FRAME = 8   # verts_aug, v_tng_aug, msdf_aug, verts_wt, v_tng_wt, msdf_wt, faces_aug, faces_wt


class _ExtractFn(torch.autograd.Function):
    """A batch of frames as ONE autograd node.

    forward : (spec, tets_i32, *unique input tensors) -> 8 tensors per frame (6 float + 2 index outputs)
    backward: dense gradients for every input tensor that needs one; a tensor shared by several frames (sdf / msdf of a
              batch of video frames, everything but msdf for the cloth / body pair) receives the sum over the frames.

    Differentiable outputs: verts_aug, msdf (augmented, stop-grad coefficients), vertices_watertight, msdf_watertight.
    v_tng_* are returned for API parity but are not differentiated (the reference's own training never consumes
    them, hmsdf.py:454,548); asking for their gradient raises instead of silently returning zeros.
    """

    @staticmethod
    def forward(ctx, spec, tets_i32, *tensors):
        frame_ids, watertight_template, lanes = spec      # frame_ids: [(pos_idx, sdf_idx, msdf_idx, negate)]
        need = ctx.needs_input_grad[2:]
        any_grad = any(need)
        gbufs: List[Optional[torch.Tensor]] = [None] * len(tensors)
        zero = None
        if any_grad:
            # dense gradient buffers of the coming backward call: allocated here, zero-filled by the tail of the
            # forward call of the first frame that uses them (HBM is idle behind the latency-bound surface kernels)
            zero, zeroed = [], set()
            for (pi, si, mi, negate) in frame_ids:
                z = []
                for idx, wanted in ((pi, True), (si, True), (mi, need[mi] and not negate)):
                    if not wanted:
                        z.append(None)
                        continue
                    if gbufs[idx] is None:
                        gbufs[idx] = torch.empty_like(tensors[idx])
                    if idx in zeroed:
                        z.append(None)
                    else:
                        zeroed.add(idx)
                        z.append(gbufs[idx].data_ptr())
                zero.append(tuple(z))
        frames = [(tensors[pi], tensors[si], tensors[mi], negate) for (pi, si, mi, negate) in frame_ids]
        results, launches = forward_frames_raw(frames, tets_i32, watertight_template, lanes, zero)
        r0 = results[0]
        ctx.save_for_backward(*tensors, r0.tape, r0.fslab)   # the slabs hold the tape and verts_wt / msdf_wt of all frames
        ctx.meta = (frame_ids, tets_i32.shape[0], [(r.n_verts, r.n_tri, r.n_quad, r.tape_ptrs, r.verts_wt.data_ptr(),
                                                      r.msdf_wt.data_ptr()) for r in results], lanes)
        ctx.gbufs = gbufs if any_grad else None
        ctx.set_materialize_grads(False)
        flat = []
        nondiff = []
        for r in results:
            flat += [r.verts_aug, r.v_tng_aug, r.msdf_aug, r.verts_wt, r.v_tng_wt, r.msdf_wt, r.faces_aug, r.faces_wt]
            nondiff += [r.faces_aug, r.faces_wt]
        ctx.mark_non_differentiable(*nondiff)
        _ExtractFn.last_counts = [r.counts for r in results]
        _ExtractFn.last_launches = launches
        return tuple(flat)

    @staticmethod
    def backward(ctx, *grads):
        frame_ids, n_tets, metas, lanes = ctx.meta
        saved = ctx.saved_tensors
        tensors = saved[:-2]
        need = ctx.needs_input_grad[2:]
        gbufs, ctx.gbufs = ctx.gbufs, None       # the pre-zeroed buffers serve ONE backward pass
        out = backward_frames_raw(tensors, frame_ids, n_tets, metas, grads, need, gbufs, lanes)
        return (None, None) + tuple(out)


def backward_frames_raw(tensors, frame_ids, n_tets, metas, grads, need, gbufs, lanes):
    """Adjoints of a batch in ONE library call.  grads: 8 upstream gradients per frame (None = zero).
    Returns one dense gradient (or None) per input tensor."""
    L = _cabi.lib()
    dev = tensors[0].device
    n_grid = tensors[frame_ids[0][0]].shape[0]
    plan = _plan_for(dev, n_tets, n_grid)
    B = len(frame_ids)
    plan.ensure(B, 1)
    ba = plan.bargs[B]
    with torch.cuda.device(dev):
        prezeroed = gbufs is not None
        if not prezeroed:  # second backward through the same node (retain_graph): fresh zero-filled buffers
            gbufs = [None] * len(tensors)
            for (pi, si, mi, negate) in frame_ids:
                for idx, wanted in ((pi, True), (si, True), (mi, need[mi] and not negate)):
                    if wanted and gbufs[idx] is None:
                        gbufs[idx] = torch.zeros_like(tensors[idx])
        keep = []
        n = 0
        for i, (pi, si, mi, negate) in enumerate(frame_ids):
            g = grads[_OUTS_PER_FRAME * i:_OUTS_PER_FRAME * (i + 1)]
            if g[1] is not None or g[4] is not None:
                raise NotImplementedError(
                    "gradients through v_tng (vertex tangents) are not implemented; D3-Human never uses them "
                    "(hmsdf.py:454,548 drop v_tng). Detach v_tng before using it in a loss.")
            if g[0] is None and g[2] is None and g[3] is None and g[5] is None:
                continue  # nothing flows into this frame
            n_verts, n_tri, n_quad, tape_ptrs, p_vwt, p_mwt = metas[i]
            va = n_verts + 3 * n_tri + 4 * n_quad
            ptrs = []
            for t, shape in ((g[0], (va, 3)), (g[2], (va,)), (g[3], (n_verts, 3)), (g[5], (n_verts,))):
                if t is None:
                    ptrs.append(None)
                    continue
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.contiguous().float()
                assert tuple(t.shape) == shape, (tuple(t.shape), shape)
                keep.append(t)
                ptrs.append(t.data_ptr())
            b = ba[n]
            n += 1
            b.pos, b.sdf, b.msdf, b.n_grid = tensors[pi].data_ptr(), tensors[si].data_ptr(), tensors[mi].data_ptr(), n_grid
            b.msdf_negate = int(negate)
            b.grads_prezeroed = 1
            b.tape_edges, b.tape_corners, b.tape_slots, b.tape_runs = tape_ptrs
            b.verts_wt, b.msdf_wt = p_vwt, p_mwt
            b.n_verts, b.n_tri_tets, b.n_quad_tets = n_verts, n_tri, n_quad
            b.g_verts_aug, b.g_msdf_aug, b.g_verts_wt, b.g_msdf_wt = ptrs
            b.g_pos, b.g_sdf = gbufs[pi].data_ptr(), gbufs[si].data_ptr()
            gm = gbufs[mi] if (need[mi] and not negate) else None
            b.g_msdf = gm.data_ptr() if gm is not None else None
            b.workspace, b.workspace_bytes = None, 0
        if n:
            _cabi.check(L.d3h_extract_backward_batch(ba, n, max(1, min(lanes, n)),
                                                     torch.cuda.current_stream(dev).cuda_stream),
                        "d3h_extract_backward_batch")
    return [gbufs[i] if need[i] else None for i in range(len(tensors))]


_ExtractFn.last_counts = None
_ExtractFn.last_launches = 0


def last_counts() -> Optional[Dict[str, int]]:
    """Sizes of the most recent extraction (Fv, T1, T2, P, V, Va, Fw, Fa, bucket sizes); frame 0 of a batch."""
    c = _ExtractFn.last_counts
    return c[0] if c else None


def last_counts_frames() -> Optional[List[Dict[str, int]]]:
    return _ExtractFn.last_counts


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (the kernels use 16-byte vector loads); a view at an odd offset is cloned."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


def _prep_pos(pos_nx3):
    if not pos_nx3.is_cuda:
        raise RuntimeError("d3human-code_b200 has no CPU path: inputs must live on a CUDA device "
                           "(the reference hard-codes device='cuda' as well, gshell_tets.py:108)")
    if pos_nx3.dim() != 2 or pos_nx3.shape[1] != 3:
        raise ValueError(f"pos_nx3 must have shape (N,3), got {tuple(pos_nx3.shape)}")
    return _aligned(pos_nx3.float())


def _prep_field(f, n_grid):
    f = f.float().reshape(-1)       # gshell_tets.py:254 (.float()); (N,1) from the SDF MLP or (N,)
    if f.shape[0] != n_grid:
        raise ValueError("sdf_n / msdf_n must have one value per grid vertex")
    return _aligned(f)


def _pack_result(r8, output_watertight_template):
    verts_aug, v_tng_aug, msdf_aug, verts_wt, v_tng_wt, msdf_wt, faces_aug, faces_wt = r8
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


def extract(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_negate: bool = False, output_watertight_template: bool = True):
    """Shared body of GShell_Tets.__call__ / hmSDF_Tets.__call__: returns the reference's 6-tuple."""
    pos = _prep_pos(pos_nx3)
    n_grid = pos.shape[0]
    sdf, msdf = _prep_field(sdf_n, n_grid), _prep_field(msdf_n, n_grid)
    tets = packed_tets(tet_fx4, n_grid)
    spec = (((0, 1, 2, bool(msdf_negate)),), bool(output_watertight_template), 1)
    return _pack_result(_ExtractFn.apply(spec, tets, pos, sdf, msdf), output_watertight_template)


def extract_frames(pos_frames, sdf_n, msdf_n, tet_fx4, types=None, output_watertight_template: bool = True,
                   lanes: int = DEFAULT_LANES):
    """A batch of extractions on the same tet grid in one autograd node and one library call per direction.

    No counterpart in the reference, which would loop over the frames (BASELINE.json configs[3]: a batch of video frames
    per step with per-frame tet-vertex offsets; also the cloth / body pair of train.py:1040-1047).

    pos_frames : (B,N,3) tensor or a sequence of (N,3) tensors -- per-frame deformed grid vertices
    sdf_n, msdf_n : one tensor shared by all frames ((N,) or (N,1)), or a sequence with one tensor per frame
    types : None (GShell_Tets semantics), one of "cloth" / "body" for all frames, or a sequence per frame
            (hmSDF_Tets semantics: "body" uses -msdf and, like the reference, does not back-propagate into msdf_n)
    Returns a list with the reference's 6-tuple `(verts, faces, None, None, v_tng, extra)` for every frame.  Gradients of
    shared tensors are summed over the frames.
    """
    if torch.is_tensor(pos_frames):
        if pos_frames.dim() != 3:
            raise ValueError(f"pos_frames must be (B,N,3), got {tuple(pos_frames.shape)}")
        pos_list = list(pos_frames.unbind(0))
    else:
        pos_list = list(pos_frames)
    B = len(pos_list)
    if B == 0:
        return []
    tensors: List[torch.Tensor] = []
    index: Dict[int, int] = {}

    def intern(src, prep):
        k = id(src)
        if k not in index:
            index[k] = len(tensors)
            tensors.append(prep(src))
        return index[k]

    pos_ids = [intern(p, _prep_pos) for p in pos_list]
    n_grid = tensors[pos_ids[0]].shape[0]
    prep_f = lambda f: _prep_field(f, n_grid)  # noqa: E731
    sdf_list = [sdf_n] * B if torch.is_tensor(sdf_n) else list(sdf_n)
    msdf_list = [msdf_n] * B if torch.is_tensor(msdf_n) else list(msdf_n)
    type_list = [types] * B if (types is None or isinstance(types, str)) else list(types)
    if not (len(sdf_list) == len(msdf_list) == len(type_list) == B):
        raise ValueError("sdf_n / msdf_n / types must be shared or have one entry per frame")
    frame_ids = tuple((pos_ids[i], intern(sdf_list[i], prep_f), intern(msdf_list[i], prep_f), type_list[i] == "body")
                      for i in range(B))
    for t in tensors:
        if t.shape[0] != n_grid:
            raise ValueError("all frames must live on the same tet grid (same N)")
    tets = packed_tets(tet_fx4, n_grid)
    spec = (frame_ids, bool(output_watertight_template), int(lanes))
    flat = _ExtractFn.apply(spec, tets, *tensors)
    return [_pack_result(flat[_OUTS_PER_FRAME * i:_OUTS_PER_FRAME * (i + 1)], output_watertight_template)
            for i in range(B)]
