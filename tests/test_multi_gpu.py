"""Real multi-rank runs (NCCL, one process per GPU): need >= 2 GPUs, skipped otherwise (`gpurun --gpus 2 -- python -m
pytest tests/test_multi_gpu.py -m gpu`).  Parity target = the single-GPU result (the reference has no multi-GPU path)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from d3human_code_b200 import grids, sharding as S
    from d3human_code_b200.extract import extract, extract_frames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        res = 32
        pos, tets = grids.kuhn_grid(res)
        sdf, msdf = grids.sphere_plane_field(pos)
        tt = torch.tensor(tets, device=dev)
        # ---- tet ranges: every rank ends up with the whole mesh, bit-identical to the single call ----
        tp = torch.tensor(pos, device=dev, requires_grad=True)
        ts = torch.tensor(sdf, device=dev, requires_grad=True)
        tm = torch.tensor(msdf, device=dev, requires_grad=True)
        verts, faces, _, _, _, extra = S.extract_tet_sharded(tp, ts, tm, tt)
        (verts.square().sum() + extra["msdf"].sum()).backward()
        v1, f1, _, _, _, e1 = extract(tp.detach(), ts.detach(), tm.detach(), tt)
        assert torch.equal(verts.detach(), v1) and torch.equal(faces, f1)
        assert torch.equal(extra["faces_watertight"], e1["faces_watertight"])
        torch.save((verts.detach().cpu(), faces.cpu(), tp.grad.cpu()), os.path.join(out_dir, f"tet_{rank}.pt"))
        # ---- frames: 4 frames over the ranks, shared sdf / msdf gradients reduced (dense == sparse) ----
        B = max(4, world)
        mine = S.frame_slice(B, world, rank)
        pos_b = torch.tensor(np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in mine]), device=dev)
        grads = []
        for sparse in (False, True):
            ts2 = torch.tensor(sdf, device=dev, requires_grad=True)
            tm2 = torch.tensor(msdf, device=dev, requires_grad=True)
            outs = extract_frames(pos_b, ts2, tm2, tt, types="cloth", lanes=2)
            sum(o[0].square().sum() + o[5]["msdf"].sum() for o in outs).backward()
            S.allreduce_shared_grads([ts2.grad, tm2.grad], sparse=sparse)
            grads.append((ts2.grad.clone(), tm2.grad.clone()))
        for a, b in zip(*grads):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * float(a.abs().max()))
        torch.save(tuple(g.cpu() for g in grads[0]), os.path.join(out_dir, f"frames_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_tet_ranges_and_frames(tmp_path, world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    import torch.multiprocessing as mp
    from d3human_code_b200 import grids
    from d3human_code_b200.extract import extract_frames
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    t0 = torch.load(tmp_path / "tet_0.pt")
    f0 = torch.load(tmp_path / "frames_0.pt")
    for r in range(1, world):
        t1 = torch.load(tmp_path / f"tet_{r}.pt")
        assert torch.equal(t0[0], t1[0]) and torch.equal(t0[1], t1[1])
        assert torch.allclose(t0[2], t1[2], rtol=1e-5, atol=1e-6 * float(t0[2].abs().max()))
        f1 = torch.load(tmp_path / f"frames_{r}.pt")
        assert torch.equal(f0[0], f1[0]) and torch.equal(f0[1], f1[1])
    # the reduced gradient equals the gradient of all the frames on one GPU
    dev = torch.device("cuda:0")
    res = 32
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = grids.sphere_plane_field(pos)
    pos_b = torch.tensor(np.stack([pos + grids.frame_offsets(pos.shape[0], res, f) for f in range(max(4, world))]), device=dev)
    ts = torch.tensor(sdf, device=dev, requires_grad=True)
    tm = torch.tensor(msdf, device=dev, requires_grad=True)
    outs = extract_frames(pos_b, ts, tm, torch.tensor(tets, device=dev), types="cloth")
    sum(o[0].square().sum() + o[5]["msdf"].sum() for o in outs).backward()
    assert torch.allclose(ts.grad.cpu(), f0[0], rtol=1e-4, atol=1e-5 * float(f0[0].abs().max()))
    assert torch.allclose(tm.grad.cpu(), f0[1], rtol=1e-4, atol=1e-5 * float(f0[1].abs().max()))
