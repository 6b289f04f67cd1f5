"""Multi-GPU host logic on CPU: world_size-2 gloo process group for the collectives around the path
(ragged record all-gather in rank order, dense / sparse reduction of the shared gradients), plus the partition rules.
The GPU side of the tet-range path (classify-range -> records -> replicated surface stages) is covered on one GPU with
virtual ranks in tests/test_cuda_parity.py; `gpurun --gpus 2` runs tests/test_multi_gpu.py on real ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from d3human_code_b200 import sharding as S


def test_frame_slice_partitions_every_frame_once():
    for n in (0, 1, 2, 7, 16, 17):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                sl = S.frame_slice(n, world, r)
                seen += list(sl)
                assert len(sl) in (n // world, n // world + 1)
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        S.frame_slice(4, 2, 2)


def test_tet_ranges_are_contiguous_tile_aligned_and_cover():
    for n_tets in (0, 1, 8191, 8192, 8193, 1_572_864, 12_582_912, 100_663_296 + 5):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = S.tet_range(n_tets, world, r)
                assert lo == prev and lo <= hi <= n_tets
                if r < world - 1:
                    assert hi % 8192 == 0 or hi == n_tets
                prev = hi
            assert prev == n_tets


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- ragged record gather: rank r holds n_r records whose first word encodes (rank, row) ----
        for counts in ([5, 3], [0, 4], [0, 0], [7, 0]):
            n = counts[rank]
            cap = 6 if rank == 0 else 4          # capacity smaller than the longest list on the other rank is fine
            cap = max(cap, n)
            local = torch.zeros((cap, S.RECORD_WORDS), dtype=torch.int32)
            local[:n, 0] = rank * 1000 + torch.arange(n, dtype=torch.int32)
            local[:n, 4] = 1 + rank
            merged, got = S.gather_records(local, n)
            assert got == counts
            want = [r * 1000 + i for r in range(world) for i in range(counts[r])]
            assert merged[:, 0].tolist() == want, (merged[:, 0].tolist(), want)
            assert merged.shape == (sum(counts), S.RECORD_WORDS)
        # ---- shared gradient reduction: dense and sparse give the same sums on every rank ----
        g = torch.Generator().manual_seed(rank)
        n_grid = 5000
        base = torch.zeros(n_grid)
        idx = torch.randperm(n_grid, generator=g)[:40 + 10 * rank]
        base[idx] = torch.randn(idx.numel(), generator=g)
        dense, sparse, forced = base.clone(), base.clone(), base.clone()
        S.allreduce_shared_grads([dense, None], sparse=False)
        S.allreduce_shared_grads([sparse], sparse=True)
        S.allreduce_shared_grads([forced], sparse=True, sparse_threshold=0.0)   # density check falls back to dense
        assert torch.allclose(dense, sparse, rtol=0, atol=1e-6) and torch.equal(dense, forced)
        torch.save(dense, os.path.join(out_dir, f"dense_{rank}.pt"))
        mats = torch.zeros((n_grid // 100, 3))
        mats[rank] = 1.0 + rank
        S.allreduce_shared_grads([mats], sparse=True)
        assert mats[0].tolist() == [1.0] * 3 and mats[1].tolist() == [2.0] * 3 and float(mats[2:].abs().sum()) == 0.0
    finally:
        dist.destroy_process_group()


def test_collectives_world_size_2_gloo(tmp_path):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    a, b = (torch.load(tmp_path / f"dense_{r}.pt") for r in range(world))
    assert torch.equal(a, b)
    assert int((a != 0).sum()) > 40
