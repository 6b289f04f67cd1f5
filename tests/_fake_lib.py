"""A CPU stand-in for libd3h_tets.so, for testing the Python HOST logic without a GPU (tests only).

It implements the C-ABI entry points the host calls (same names, same argument blocks read from the same raw
addresses) and computes the results with the numpy oracle, honouring the capacity / overflow contract of the real
library: outputs are truncated to the caller's capacities, true sizes come back in d3h_counts, a too-small record
capacity skips the surface stages.  The product never sees this module; `tests/test_host_fake_lib.py` installs it with
monkeypatch in place of `_cabi.lib()` and runs the real `extract.py` code paths on CPU tensors.
"""
import ctypes as C

import numpy as np

from d3human_code_b200 import _cabi
from oracle import gshell_oracle as O


def _arr(ptr, n, ctype):
    if not ptr or n <= 0:
        return np.zeros(0, dtype=np.dtype(ctype))
    return np.ctypeslib.as_array((ctype * int(n)).from_address(int(ptr)))


class FakeLib:
    def __init__(self):
        self.tapes = {}          # tape_edges pointer -> oracle forward dict (the "tape" of the fake)
        self.tapes_pair = None   # forward dict of the second extraction of the last fused pair
        self.forward_calls = 0   # frames
        self.batch_calls = 0
        self.backward_calls = 0
        self.joins = 0
        self.error = b""

    # ---- trivial entry points -------------------------------------------------------------------------------------
    def d3h_version(self):
        return _cabi.VERSION

    def d3h_last_error_string(self):
        return self.error

    def d3h_workspace_bytes(self, n_tets, n_grid, cap):
        return self.d3h_workspace_bytes_static(n_tets, n_grid, cap, 0)

    def d3h_workspace_bytes_static(self, n_tets, n_grid, cap, n_edges):
        return 4096 + 64 * cap + n_edges // 4      # monotone in the capacities like the real one

    def d3h_gather_rows(self, ids, n_ids, src, n_rows, width, out, stream):
        i = _arr(ids, n_ids, C.c_int32)
        g = _arr(src, n_rows * width, C.c_float).reshape(n_rows, width)
        _arr(out, n_ids * width, C.c_float).reshape(n_ids, width)[:] = g[i]
        return 0

    def d3h_lanes_join(self, stream):
        self.joins += 1
        return 0

    def d3h_wait_counts(self, ptr, seq, timeout_us):
        c = _cabi.Counts.from_address(int(ptr))
        if c.seq != seq:
            self.error = b"fake: counts not published"
            return _cabi.D3H_E_TIMEOUT
        return 0

    # ---- forward --------------------------------------------------------------------------------------------------
    def d3h_extract_forward_batch(self, ptr, n_frames, lanes, stream):
        return self.d3h_extract_forward_batch_nojoin(ptr, n_frames, lanes, stream)

    def d3h_extract_forward(self, ptr, stream):          # the single-call path (single.py)
        self.single_calls = getattr(self, "single_calls", 0) + 1
        return self._forward(_cabi.ForwardArgs.from_address(int(ptr)))

    def d3h_extract_backward(self, ptr, stream):
        return self.d3h_extract_backward_batch(ptr, 1, 1, stream)

    def d3h_extract_forward_batch_nojoin(self, ptr, n_frames, lanes, stream):
        self.batch_calls += 1
        size = C.sizeof(_cabi.ForwardArgs)
        seen = {}
        for i in range(n_frames):
            a = _cabi.ForwardArgs.from_address(int(ptr) + i * size)
            lane = i % max(1, min(lanes, n_frames))
            if seen.setdefault(a.workspace, lane) != lane:
                self.error = b"fake: frames on different lanes share a workspace"
                return _cabi.D3H_E_BADARG
            rc = self._forward(a)
            if rc:
                return rc
        return 0

    def _forward(self, a, pair_of=None):
        if pair_of is None and a.pair_counts_host:
            # fused cloth / body pair: the second extraction is the same call with the sign flipped and the pair_* outputs
            rc = self._forward(a, pair_of=False)
            if rc:
                return rc
            b = _cabi.ForwardArgs.from_buffer_copy(bytes(a))
            b.msdf_negate = 0 if a.msdf_negate else 1
            for name in ("verts_aug", "v_tng_aug", "msdf_aug", "faces_aug", "verts_wt", "v_tng_wt", "msdf_wt", "faces_wt", "vacc"):
                setattr(b, name, getattr(a, "pair_" + name))
            b.counts_host, b.seq = a.pair_counts_host, a.pair_seq
            b.zero_g_pos = b.zero_g_sdf = b.zero_g_msdf = None
            b.pair_counts_host = None
            rc = self._forward(b, pair_of=True)
            # one tape for both: the backward blocks of frame 1 point at frame 0's tape, but carry the negated sign
            self.tapes[(int(a.tape_edges), 1)] = self.tapes_pair
            return rc
        self.forward_calls += 1
        n, f = a.n_grid, a.n_tets
        if (a.sdf | a.msdf) & 15 or a.workspace & 63 or a.tets & 15:   # (CPU allocations are 64-byte aligned)
            self.error = b"fake: misaligned pointer"
            return _cabi.D3H_E_BADARG
        pos = _arr(a.pos, 3 * n, C.c_float).reshape(n, 3)
        sdf = _arr(a.sdf, n, C.c_float)
        msdf = _arr(a.msdf, n, C.c_float)
        tets = _arr(a.tets, 4 * f, C.c_int32).reshape(f, 4)
        fwd = O.extract_forward(pos.copy(), sdf.copy(), msdf.copy(), tets, -1 if a.msdf_negate else 1,
                                bool(a.watertight_template))
        for zp, ln in ((a.zero_g_pos, 3 * n), (a.zero_g_sdf, n), (a.zero_g_msdf, n)):
            if zp:
                _arr(zp, ln, C.c_float)[:] = 0.0
        fv, t1, t2 = fwd["fv"], fwd["t1"], fwd["t2"]
        c = _cabi.Counts.from_address(int(a.counts_host))
        c.n_valid_tets, c.n_tri_tets, c.n_quad_tets, c.n_corners = fv, t1, t2, 3 * t1 + 4 * t2
        c.bad_index = 0
        if fv > a.cap_valid_tets:      # record buffer too small: the surface stages are skipped
            c.n_verts = c.n_faces_aug = 0
            for k in range(6):
                c.bucket_polys[k] = 0
            c.overflow = 1
            c.seq = a.seq
            return 0
        v = fwd["n_verts_watertight"]
        va, fa, fw = fwd["verts_aug"].shape[0], fwd["faces_aug"].shape[0], fwd["faces_watertight"].shape[0]

        def put(ptr, cap_rows, src, ctype, width):
            rows = min(cap_rows, src.shape[0])
            if rows > 0:
                _arr(ptr, rows * width, ctype).reshape(rows, width)[:] = src[:rows].reshape(rows, width)

        put(a.verts_aug, a.cap_verts_aug, fwd["verts_aug"], C.c_float, 3)
        put(a.v_tng_aug, a.cap_verts_aug, fwd["v_tng_aug"], C.c_float, 3)
        put(a.msdf_aug, a.cap_verts_aug, fwd["msdf"], C.c_float, 1)
        put(a.faces_aug, a.cap_faces_aug, fwd["faces_aug"], C.c_int64, 3)
        put(a.verts_wt, a.cap_verts, fwd["vertices_watertight"], C.c_float, 3)
        put(a.v_tng_wt, a.cap_verts, fwd["v_tng_watertight"], C.c_float, 3)
        put(a.msdf_wt, a.cap_verts, fwd["msdf_watertight"], C.c_float, 1)
        put(a.faces_wt, a.cap_faces_wt, fwd["faces_watertight"], C.c_int64, 3)
        put(a.tape_edges, a.cap_verts, np.stack([fwd["edge_a"], fwd["edge_b"]], 1).astype(np.int32), C.c_int32, 2)
        put(a.tape_corners, 4 * a.cap_valid_tets, fwd["corners"].astype(np.int32), C.c_int32, 1)
        if a.edge_off:
            if not a.vacc or not a.edge_ab or a.n_edges <= 0:
                self.error = b"fake: static edge table without vacc / edge_ab"
                return _cabi.D3H_E_BADARG
            _arr(a.vacc, 8 * min(a.cap_verts, v), C.c_float)[:] = 0.0
        if pair_of:
            self.tapes_pair = fwd                 # same tape address as the first extraction of the pair
        else:
            self.tapes[int(a.tape_edges)] = fwd
            self.tapes[("corners", int(a.tape_corners))] = fwd      # the tangent branch finds the call by its corner array
        c.n_verts, c.n_faces_aug = v, fa
        per = (1, 2, 1, 2, 3, 4)
        for k in range(6):
            c.bucket_polys[k] = int(fwd["bucket_counts"][k]) // per[k]
        c.overflow = 0
        c.seq = a.seq
        return 0

    # ---- tet-range sharding ----------------------------------------------------------------------------------------
    def d3h_classify_range(self, ptr, records_out, cap_records, counts_dev_out, stream):
        a = _cabi.ForwardArgs.from_address(int(ptr))
        n, f = a.n_grid, a.n_tets
        sdf = _arr(a.sdf, n, C.c_float)
        msdf = _arr(a.msdf, n, C.c_float)
        tets = _arr(a.tets, 4 * f, C.c_int32).reshape(f, 4)
        occ = sdf > 0
        m = -msdf if a.msdf_negate else msdf
        lo, hi = a.tet_begin, a.tet_end
        sub = tets[lo:hi].astype(np.int64)
        o = occ[sub]
        code = (o * np.array([1, 2, 4, 8])).sum(1)
        valid = (code != 0) & (code != 15)
        if not a.watertight_template:
            valid &= (m[sub] > 0).any(1)
        ids = np.nonzero(valid)[0]
        cls2 = np.isin(code[ids], (3, 5, 6, 9, 10, 12))
        rank1, rank2 = np.cumsum(~cls2) - 1, np.cumsum(cls2) - 1
        n_valid = ids.shape[0]
        rec = np.zeros((n_valid, 8), np.int32)
        rec[:, :4] = sub[ids]
        rec[:, 4] = code[ids]
        rec[:, 5] = np.where(cls2, rank2, rank1)
        rec[:, 6] = lo + ids
        rec[:, 7] = np.where(cls2, rank1 + 1, rank2 + 1)
        fits = n_valid <= cap_records
        if fits and n_valid:
            _arr(records_out, 8 * n_valid, C.c_int32).reshape(-1, 8)[:] = rec
        for p in (counts_dev_out, a.counts_host):
            if not p:
                continue
            c = _cabi.Counts.from_address(int(p))
            c.n_valid_tets, c.n_tri_tets, c.n_quad_tets = n_valid, int((~cls2).sum()), int(cls2.sum())
            c.n_corners = 3 * c.n_tri_tets + 4 * c.n_quad_tets
            c.overflow = 0 if fits else 1
            c.seq = a.seq
        return 0

    def d3h_extract_from_records(self, ptr, records, n_tri, n_quad, stream):
        a = _cabi.ForwardArgs.from_address(int(ptr))
        n_rec = n_tri + n_quad
        if n_rec > a.cap_valid_tets:
            self.error = b"fake: records do not fit cap_valid_tets"
            return _cabi.D3H_E_BADARG
        rec = _arr(records, 8 * n_rec, C.c_int32).reshape(-1, 8)
        if n_rec and not np.all(np.diff(rec[:, 6]) > 0):
            self.error = b"fake: records are not in global tet order"
            return _cabi.D3H_E_BADARG
        # the surface stages only ever look at the valid tets: run the oracle on the full tet array the records came
        # from (same result, and the UV atlas keeps the size of the full grid) after checking the records against it
        f = a.n_tets
        tets = _arr(a.tets, 4 * f, C.c_int32).reshape(f, 4)
        if n_rec and not np.array_equal(tets[rec[:, 6]], rec[:, :4]):
            self.error = b"fake: record vertices do not match their tet ids"
            return _cabi.D3H_E_BADARG
        self.from_records_calls = getattr(self, "from_records_calls", 0) + 1
        saved = a.edge_off
        a.edge_off = None             # the sharded stages always take the general path
        rc = self._forward(a)
        a.edge_off = saved
        c = _cabi.Counts.from_address(int(a.counts_host))
        if rc == 0 and c.n_valid_tets != n_rec:
            self.error = b"fake: gathered records are not the valid tets of the grid"
            return _cabi.D3H_E_BADARG
        return rc

    # ---- tangent branch (include/d3h_tets.h: d3h_tangent_backward) -----------------------------------------------------
    def d3h_tangent_backward(self, ptr, stream):
        a = _cabi.TangentBackwardArgs.from_address(int(ptr))
        fwd = self.tapes.get(("corners", int(a.tape_corners)))
        if fwd is None:
            self.error = b"fake: unknown tape (tangent branch)"
            return _cabi.D3H_E_BADARG
        if fwd["faces_watertight"].shape[0] == 3:
            self.error = b"fake: a watertight mesh of exactly three faces is not supported"
            return _cabi.D3H_E_BADARG
        v, va = fwd["n_verts_watertight"], fwd["verts_aug"].shape[0]
        g_aug = _arr(a.g_tng_aug, 3 * va, C.c_float).reshape(-1, 3) if a.g_tng_aug else None
        g_wt = _arr(a.g_tng_wt, 3 * v, C.c_float).reshape(-1, 3) if a.g_tng_wt else None
        g_vert, g_mv = O.tangent_backward(fwd, g_aug, g_wt)
        _arr(a.g_verts, 3 * v, C.c_float)[:] = g_vert.astype(np.float32).reshape(-1)
        _arr(a.g_mvert, v, C.c_float)[:] = g_mv.astype(np.float32)
        return 0

    # ---- backward -------------------------------------------------------------------------------------------------
    def d3h_extract_backward_batch(self, ptr, n_frames, lanes, stream):
        size = C.sizeof(_cabi.BackwardArgs)
        for i in range(n_frames):
            b = _cabi.BackwardArgs.from_address(int(ptr) + i * size)
            self.backward_calls += 1
            fwd = self.tapes.get(int(b.tape_edges))
            pair_fwd = self.tapes.get((int(b.tape_edges), 1))
            if pair_fwd is not None and fwd is not None and (-1 if b.msdf_negate else 1) != fwd["_msdf_sign"]:
                fwd = pair_fwd                    # second extraction of a fused pair
            if fwd is None:
                self.error = b"fake: unknown tape"
                return _cabi.D3H_E_BADARG
            if not b.tape_slots and not b.vacc:
                self.error = b"fake: neither corner lists nor vacc"
                return _cabi.D3H_E_BADARG
            v = fwd["n_verts_watertight"]
            va = fwd["verts_aug"].shape[0]
            if (b.n_verts, b.n_tri_tets, b.n_quad_tets) != (v, fwd["t1"], fwd["t2"]):
                self.error = b"fake: sizes of the backward block do not match the forward call"
                return _cabi.D3H_E_BADARG
            gva = _arr(b.g_verts_aug, 3 * va, C.c_float).reshape(-1, 3) if b.g_verts_aug else None
            gma = _arr(b.g_msdf_aug, va, C.c_float).copy() if b.g_msdf_aug else None
            if b.g_msdf_boundary:
                if gma is None:
                    gma = np.zeros(va, np.float32)
                gma[v:] += _arr(b.g_msdf_boundary, va - v, C.c_float)
            gvw = _arr(b.g_verts_wt, 3 * v, C.c_float).reshape(-1, 3) if b.g_verts_wt else None
            gmw = _arr(b.g_msdf_wt, v, C.c_float) if b.g_msdf_wt else None
            if b.g_verts_tng:     # through the tangent branch: added to the per-vertex gradients
                extra_v = _arr(b.g_verts_tng, 3 * v, C.c_float).reshape(-1, 3)
                gvw = extra_v.copy() if gvw is None else gvw + extra_v
            gmx = _arr(b.g_mvert_tng, v, C.c_float) if b.g_mvert_tng else None
            g_pos, g_sdf, g_msdf = O.extract_backward(fwd, gva, gma, gvw, gmw, None, None, gmx)
            n = b.n_grid
            outs = ((b.g_pos, 3 * n, g_pos), (b.g_sdf, n, g_sdf), (b.g_msdf, n, g_msdf))
            for p, ln, g in outs:
                if not p:
                    continue
                dst = _arr(p, ln, C.c_float)
                if not b.grads_prezeroed:
                    dst[:] = 0.0
                if g is not None:
                    dst += np.asarray(g, np.float32).reshape(-1)     # shared buffers accumulate over the frames
        return 0
