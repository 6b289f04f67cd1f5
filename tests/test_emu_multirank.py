"""The multi-rank paths end to end on CPU: two gloo ranks, each with the emulated kernels behind the real C ABI
(tests/emu), run what tests/test_multi_gpu.py runs on two GPUs over NCCL: tet-range sharding (every rank classifies its
range, the records are all-gathered in rank order, the surface stages run replicated) and frame sharding with the
reduction of the shared gradients.  Parity target: the single-rank result."""
import contextlib
import os
import socket
import sys

import numpy as np
import torch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Setter:
    """monkeypatch-like setter for the spawned processes (which have no fixture and never need to undo)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def _patch_for_cpu(mp=_Setter):
    """What the `dev` fixture of tests/test_emu_parity.py does; `mp` = pytest's monkeypatch in the parent process."""
    from d3human_code_b200 import _cabi
    from d3human_code_b200 import extract as E
    from tests import test_emu_parity as T
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    mp.setattr(_cabi, "LIB_PATH", build_emu.build())
    mp.setattr(_cabi, "_lib", None)
    mp.setattr(E, "_check_cuda", lambda t: None)
    mp.setattr(E, "packed_tets", T._packed_tets_cpu)
    mp.setattr(torch.cuda, "current_stream", lambda dev=None: T._Stream())
    mp.setattr(torch.cuda, "device", lambda dev=None: contextlib.nullcontext())
    mp.setattr(torch.Tensor, "pin_memory", lambda self: self)
    orig = E._Plan.ensure

    def ensure(self, lanes, n_edges=0):
        orig(self, lanes, n_edges)
        for i, w in enumerate(self.workspaces):
            if w.data_ptr() % 256:
                big = torch.empty(self.workspace_bytes + 256, dtype=torch.uint8)
                off = (-big.data_ptr()) % 256
                self.workspaces[i] = big[off:off + self.workspace_bytes]
                self.workspace_ptrs[i] = self.workspaces[i].data_ptr()

    mp.setattr(E._Plan, "ensure", ensure)
    E.reset_plans()
    return E


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        E = _patch_for_cpu()
        from d3human_code_b200 import grids, sharding as S
        res = 12
        pos, tets = grids.kuhn_grid(res)
        sdf, msdf = grids.capsule_garment_field(pos)
        tt = torch.tensor(tets)
        # ---- tet ranges ----
        tp, ts, tm = torch.tensor(pos, requires_grad=True), torch.tensor(sdf, requires_grad=True), torch.tensor(msdf, requires_grad=True)
        verts, faces, _, _, _, extra = S.extract_tet_sharded(tp, ts, tm, tt)
        (verts.square().sum() + extra["msdf"].sum()).backward()
        v1, f1, _, _, _, e1 = E.extract(tp.detach(), ts.detach(), tm.detach(), tt)
        assert torch.equal(verts.detach(), v1) and torch.equal(faces, f1)
        assert torch.equal(extra["faces_watertight"], e1["faces_watertight"])
        lo, hi = S.tet_range(tets.shape[0], world, rank)
        assert 0 <= lo < hi <= tets.shape[0]
        torch.save((verts.detach(), faces, tp.grad, ts.grad), os.path.join(out_dir, f"tet_{rank}.pt"))
        # ---- frames ----
        B = 4
        mine = S.frame_slice(B, world, rank)
        pos_b = torch.tensor(np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in mine]))
        grads = []
        for sparse in (False, True):
            ts2, tm2 = torch.tensor(sdf, requires_grad=True), torch.tensor(msdf, requires_grad=True)
            outs = E.extract_frames(pos_b, ts2, tm2, tt, types="cloth", lanes=2)
            sum(o[0].square().sum() + o[5]["msdf"].sum() for o in outs).backward()
            S.allreduce_shared_grads([ts2.grad, tm2.grad], sparse=sparse)
            grads.append((ts2.grad.clone(), tm2.grad.clone()))
        for a, b in zip(*grads):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * float(a.abs().max()))
        torch.save(grads[0], os.path.join(out_dir, f"frames_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_with_emulated_kernels(tmp_path, monkeypatch):
    import torch.multiprocessing as mp
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    t0, t1 = (torch.load(tmp_path / f"tet_{r}.pt") for r in range(world))
    assert torch.equal(t0[0], t1[0]) and torch.equal(t0[1], t1[1])            # every rank holds the whole mesh
    assert torch.allclose(t0[2], t1[2], rtol=1e-5, atol=1e-6 * float(t0[2].abs().max()))
    f0, f1 = (torch.load(tmp_path / f"frames_{r}.pt") for r in range(world))
    assert torch.equal(f0[0], f1[0]) and torch.equal(f0[1], f1[1])            # reduced gradients identical on all ranks
    # the reduced gradient equals the gradient of all 4 frames on one rank
    E = _patch_for_cpu(monkeypatch)
    from d3human_code_b200 import grids
    res = 12
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = grids.capsule_garment_field(pos)
    pos_b = torch.tensor(np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in range(4)]))
    ts, tm = torch.tensor(sdf, requires_grad=True), torch.tensor(msdf, requires_grad=True)
    outs = E.extract_frames(pos_b, ts, tm, torch.tensor(tets), types="cloth")
    sum(o[0].square().sum() + o[5]["msdf"].sum() for o in outs).backward()
    assert torch.allclose(ts.grad, f0[0], rtol=1e-4, atol=1e-5 * float(f0[0].abs().max()))
    assert torch.allclose(tm.grad, f0[1], rtol=1e-4, atol=1e-5 * float(f0[1].abs().max()))
    E.reset_plans()
