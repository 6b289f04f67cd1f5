"""The Python host of the path (extract.py: capacity planning and regrowth, slab layout and output views, the batched
autograd node, async batches, backward argument blocks, shared-gradient accumulation) run end to end on CPU tensors
against `tests/_fake_lib.FakeLib`, a stand-in for libd3h_tets.so that answers the same C-ABI calls with the numpy
oracle.  This checks the host logic on machines without a GPU; the kernels themselves are covered by the `-m gpu`
parity tests."""
import contextlib

import numpy as np
import pytest
import torch

from d3human_code_b200 import _cabi, grids
from d3human_code_b200 import extract as E
from oracle import gshell_oracle as O
from tests import _util as U
from tests._fake_lib import FakeLib


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass


@pytest.fixture(params=["sort", "static"])
def host(request, monkeypatch):
    """extract.py wired to the fake library and to CPU tensors."""
    fake = FakeLib()
    monkeypatch.setattr(_cabi, "lib", lambda: fake)
    monkeypatch.setattr(E, "_check_cuda", lambda t: None)
    monkeypatch.setattr(E, "packed_tets", lambda t, n: t.to(torch.int32).contiguous())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: _Stream())
    monkeypatch.setattr(torch.cuda, "device", lambda dev=None: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    E.reset_plans()
    E.set_static_edges("1" if request.param == "static" else "0")
    yield fake
    E.set_static_edges("auto")
    E.reset_plans()


def _case(res=8, field="capsule"):
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = (grids.capsule_garment_field if field == "capsule" else grids.sphere_plane_field)(pos)
    return pos, sdf, msdf, tets


def _check_frame(out, fwd):
    verts, faces, uvs, uv_idx, v_tng, extra = out
    assert uvs is None and uv_idx is None
    U.assert_exact("verts_aug", verts.detach().numpy(), fwd["verts_aug"])
    U.assert_exact("faces_aug", faces.numpy(), fwd["faces_aug"])
    U.assert_exact("v_tng_aug", v_tng.detach().numpy(), fwd["v_tng_aug"])
    assert tuple(extra.keys()) == fwd["extra_keys"]
    U.assert_exact("msdf", extra["msdf"].detach().numpy(), fwd["msdf"])
    U.assert_exact("msdf_boundary", extra["msdf_boundary"].detach().numpy(), fwd["msdf_boundary"])
    if "vertices_watertight" in extra:
        assert extra["n_verts_watertight"] == fwd["n_verts_watertight"]
        U.assert_exact("vertices_watertight", extra["vertices_watertight"].detach().numpy(), fwd["vertices_watertight"])
        U.assert_exact("faces_watertight", extra["faces_watertight"].numpy(), fwd["faces_watertight"])


@pytest.mark.parametrize("typ,wt", [(None, True), ("cloth", True), ("body", True), ("body", False)])
def test_single_call_views_and_gradients(host, typ, wt):
    pos, sdf, msdf, tets = _case(8)
    tp = torch.tensor(pos, requires_grad=True)
    ts = torch.tensor(sdf[:, None].astype(np.float64), requires_grad=True)     # (N,1) float64 like an MLP output
    tm = torch.tensor(msdf, requires_grad=True)
    n0 = E.launch_counter()
    out = E.extract(tp, ts, tm, torch.tensor(tets), msdf_negate=(typ == "body"), output_watertight_template=wt)
    n1 = E.launch_counter()
    assert n1 - n0 >= E.LAUNCHES_FORWARD_STATIC + 1      # counting run(s) + the real run + the zero-fill kernel
    fwd = O.extract_forward(pos, sdf, msdf, tets, -1 if typ == "body" else 1, wt)
    _check_frame(out, fwd)
    verts, extra = out[0], out[5]
    v = fwd["n_verts_watertight"]
    assert extra["msdf_boundary"].data_ptr() == extra["msdf"][v:].data_ptr()
    rng = np.random.default_rng(0)
    gv = rng.standard_normal(fwd["verts_aug"].shape).astype(np.float32)
    gm = rng.standard_normal(fwd["msdf"].shape).astype(np.float32)
    gb = rng.standard_normal(fwd["msdf_boundary"].shape).astype(np.float32)
    ((verts * torch.tensor(gv)).sum() + (extra["msdf"] * torch.tensor(gm)).sum()
     + (extra["msdf_boundary"] * torch.tensor(gb)).sum()).backward()
    assert E.launch_counter() - n1 in (1, 2)             # adjoint (+ adjoint_poly on the static path)
    gm_total = gm.copy()
    gm_total[v:] += gb
    g_pos, g_sdf, g_msdf = O.extract_backward(fwd, gv, gm_total)
    U.assert_close_normwise("grad_pos", tp.grad.numpy(), g_pos, 1e-6)
    assert ts.grad.shape == ts.shape and ts.grad.dtype == ts.dtype
    U.assert_close_normwise("grad_sdf", ts.grad.numpy()[:, 0], g_sdf, 1e-6)
    if typ == "body":
        assert tm.grad is None
    else:
        U.assert_close_normwise("grad_msdf", tm.grad.numpy(), g_msdf, 1e-6)
    # gradients through the tangents (SURVEY A.5 optional branch): the host routes them through d3h_tangent_backward and
    # hands its two per-vertex arrays to the adjoint call
    tp.grad = ts.grad = tm.grad = None
    out = E.extract(tp, ts, tm, torch.tensor(tets), msdf_negate=(typ == "body"), output_watertight_template=wt)
    gt = np.random.default_rng(9).standard_normal(fwd["v_tng_aug"].shape).astype(np.float32)
    (out[4] * torch.tensor(gt)).sum().backward()
    g_pos, g_sdf, _ = O.extract_backward(fwd, None, None, None, None, gt, None)
    U.assert_close_normwise("grad_pos (tangents)", tp.grad.numpy(), g_pos, 1e-5)
    U.assert_close_normwise("grad_sdf (tangents)", ts.grad.numpy()[:, 0], g_sdf, 1e-5)


def test_capacity_regrowth_and_shrink(host):
    """Same grid, surface grows 10x, shrinks, vanishes: every call exact, no more than one re-run per growth."""
    pos, tets = grids.kuhn_grid(10)
    p = pos.astype(np.float64)
    tt = torch.tensor(tets)
    for radius in (0.15, 0.9, 0.3, 0.0, 0.5):
        sdf = (radius - np.linalg.norm(p, axis=-1)).astype(np.float32)
        msdf = (p[:, 1] + 0.05).astype(np.float32)
        before = host.forward_calls
        out = E.extract(torch.tensor(pos), torch.tensor(sdf), torch.tensor(msdf), tt)
        _check_frame(out, O.extract_forward(pos, sdf, msdf, tets))
        assert host.forward_calls - before <= 3      # counting run + (bounded) retry at the very first size


def test_batch_of_frames_shared_and_stacked_inputs(host):
    res, B = 8, 5
    pos, sdf, msdf, tets = _case(res)
    n = pos.shape[0]
    types = ["cloth", "body", "cloth", "cloth", "body"]
    pos_b = np.stack([pos + grids.frame_offsets(n, res, f) for f in range(B)]).astype(np.float32)
    tp = torch.tensor(pos_b, requires_grad=True)          # stacked (B,N,3): rows are not 16-byte aligned (N odd)
    ts = torch.tensor(sdf[:, None], requires_grad=True)
    tm = torch.tensor(msdf, requires_grad=True)
    outs = E.extract_frames(tp, ts, tm, torch.tensor(tets), types=types, lanes=3)
    assert host.batch_calls >= 1 and len(outs) == B and len(E.last_counts_frames()) == B
    rng = np.random.default_rng(1)
    want_pos, want_sdf, want_msdf = np.zeros_like(pos_b), np.zeros_like(sdf), np.zeros_like(msdf)
    loss = 0.0
    for i, out in enumerate(outs):
        fwd = O.extract_forward(pos_b[i], sdf, msdf, tets, -1 if types[i] == "body" else 1, True)
        _check_frame(out, fwd)
        if i == 3:
            continue                                       # a frame without upstream gradient
        gv = rng.standard_normal(fwd["verts_aug"].shape).astype(np.float32)
        gw = rng.standard_normal(fwd["vertices_watertight"].shape).astype(np.float32)
        loss = loss + (out[0] * torch.tensor(gv)).sum() + (out[5]["vertices_watertight"] * torch.tensor(gw)).sum()
        g_pos, g_sdf, g_msdf = O.extract_backward(fwd, gv, None, gw, None)
        want_pos[i] = g_pos
        want_sdf += g_sdf
        if types[i] != "body":
            want_msdf += g_msdf
    loss.backward()
    U.assert_close_normwise("grad_pos", tp.grad.numpy(), want_pos, 1e-6)
    U.assert_close_normwise("grad_sdf", ts.grad.numpy()[:, 0], want_sdf, 1e-6)
    U.assert_close_normwise("grad_msdf", tm.grad.numpy(), want_msdf, 1e-6)


def test_async_batches_in_flight_and_retain_graph(host):
    res = 8
    pos, sdf, msdf, tets = _case(res, "sphere")
    n = pos.shape[0]
    tt = torch.tensor(tets)
    ts = torch.tensor(sdf, requires_grad=True)
    tm = torch.tensor(msdf, requires_grad=True)
    groups = [torch.tensor(np.stack([pos + grids.frame_offsets(n, res, 10 * g + f) for f in range(3)]), requires_grad=True)
              for g in range(3)]
    futs = [E.extract_frames_async(pg, ts, tm, tt, types="cloth", lanes=2) for pg in groups]
    assert host.forward_calls >= 9                        # all three batches were launched before any result is read
    joins = host.joins
    want_sdf = np.zeros_like(sdf)
    for g, fut in enumerate(futs):
        outs = fut.result()
        assert fut.result() is outs                       # idempotent
        for f, out in enumerate(outs):
            fwd = O.extract_forward(groups[g][f].detach().numpy(), sdf, msdf, tets, 1, True)
            _check_frame(out, fwd)
            want_sdf += O.extract_backward(fwd, np.ones_like(fwd["verts_aug"]), None)[1]
        sum(o[0].sum() for o in outs).backward(retain_graph=(g == 0))
        if g == 0:                                        # a second pass through the same node: fresh zeroed buffers
            first = ts.grad.clone()
            sum(o[0].sum() for o in outs).backward()
            assert torch.allclose(ts.grad, 2 * first, rtol=1e-6, atol=1e-7)
            ts.grad = first
    assert host.joins > joins                             # the lanes were joined before outputs were handed out
    U.assert_close_normwise("grad_sdf", ts.grad.numpy(), want_sdf, 1e-6)
    assert E.extract_frames_async([], ts, tm, tt).result() == []


def test_input_validation_on_host(host):
    pos, sdf, msdf, tets = _case(4)
    tt = torch.tensor(tets)
    with pytest.raises(ValueError):
        E.extract(torch.tensor(pos)[:, :2], torch.tensor(sdf), torch.tensor(msdf), tt)
    with pytest.raises(ValueError):
        E.extract(torch.tensor(pos), torch.tensor(sdf)[:-1], torch.tensor(msdf), tt)
    with pytest.raises(ValueError):
        E.extract_frames(torch.tensor(pos), torch.tensor(sdf), torch.tensor(msdf), tt)        # not (B,N,3)
    with pytest.raises(ValueError):
        E.extract_frames([torch.tensor(pos)] * 2, [torch.tensor(sdf)] * 3, torch.tensor(msdf), tt)


@pytest.mark.parametrize("vr", [1, 3, 8])
def test_tet_range_sharding_host_logic(host, vr):
    """sharding.extract_tet_sharded with virtual ranks: per-range classification, concatenation of the records in range
    order, replicated surface stages -- same result as the plain call, also when the record capacity has to grow."""
    from d3human_code_b200 import sharding as S
    pos, sdf, msdf, tets = _case(10)
    tp = torch.tensor(pos, requires_grad=True)
    ts = torch.tensor(sdf, requires_grad=True)
    tm = torch.tensor(msdf, requires_grad=True)
    out = S.extract_tet_sharded(tp, ts, tm, torch.tensor(tets), virtual_ranks=vr)
    fwd = O.extract_forward(pos, sdf, msdf, tets)
    _check_frame(out, fwd)
    assert host.from_records_calls >= 1
    gv = np.ones_like(fwd["verts_aug"])
    out[0].sum().backward()
    g_pos, g_sdf, g_msdf = O.extract_backward(fwd, gv, None)
    U.assert_close_normwise("grad_pos", tp.grad.numpy(), g_pos, 1e-6)
    U.assert_close_normwise("grad_sdf", ts.grad.numpy(), g_sdf, 1e-6)
    ranges = [S.tet_range(tets.shape[0], vr, r) for r in range(vr)]
    assert ranges[0][0] == 0 and ranges[-1][1] == tets.shape[0]


@pytest.mark.parametrize("fused", [False, True])
def test_split_pair_is_cloth_plus_body(host, fused):
    """hmSDF_Tets.split: one call for the cloth / body pair of an iteration == the two reference-style calls."""
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    pos, sdf, msdf, tets = _case(8)
    tp = torch.tensor(pos, requires_grad=True)
    ts = torch.tensor(sdf[:, None], requires_grad=True)
    tm = torch.tensor(msdf, requires_grad=True)
    cloth, body = hmSDF_Tets().split(tp, ts, tm, torch.tensor(tets), fused=fused)
    fc = O.extract_forward(pos, sdf, msdf, tets, 1, True)
    fb = O.extract_forward(pos, sdf, msdf, tets, -1, True)
    _check_frame(cloth, fc)
    _check_frame(body, fb)
    (cloth[0].sum() + body[0].sum() + cloth[5]["msdf"].sum() + body[5]["msdf"].sum()).backward()
    gc = O.extract_backward(fc, np.ones_like(fc["verts_aug"]), np.ones_like(fc["msdf"]))
    gb = O.extract_backward(fb, np.ones_like(fb["verts_aug"]), np.ones_like(fb["msdf"]))
    U.assert_close_normwise("grad_pos", tp.grad.numpy(), gc[0] + gb[0], 1e-6)
    U.assert_close_normwise("grad_sdf", ts.grad.numpy()[:, 0], gc[1] + gb[1], 1e-6)
    U.assert_close_normwise("grad_msdf", tm.grad.numpy(), gc[2], 1e-6)      # cloth only


def test_count_ring_refuses_to_wrap_over_unread_batches(host):
    """batches in flight take consecutive slots of the pinned count ring (256 per grid).  A launch that would
    wrap over slots nobody has read yet must raise instead of overwriting them; reading (or dropping) the earlier futures
    frees the slots."""
    from d3human_code_b200 import extract as E
    pos, tets = grids.kuhn_grid(4)
    sdf, msdf = grids.sphere_plane_field(pos)
    B = 100
    tp = torch.tensor(np.stack([pos] * B))
    ts, tm, tt = torch.tensor(sdf), torch.tensor(msdf), torch.tensor(tets)
    f1 = E.extract_frames_async(tp, ts, tm, tt, lanes=2)
    f2 = E.extract_frames_async(tp, ts, tm, tt, lanes=2)
    with pytest.raises(RuntimeError, match="in flight"):
        E.extract_frames_async(tp, ts, tm, tt, lanes=2)
    assert len(f1.result()) == B
    f3 = E.extract_frames_async(tp, ts, tm, tt, lanes=2)      # fits again
    del f2                                                     # dropped without result(): its slots are released
    import gc
    gc.collect()
    f4 = E.extract_frames_async(tp, ts, tm, tt, lanes=2)
    assert len(f3.result()) == B and len(f4.result()) == B
    plan = E._plan_for(tp.device, tets.shape[0], pos.shape[0])
    assert plan.inflight == 0
