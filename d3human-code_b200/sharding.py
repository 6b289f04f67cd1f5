"""Multi-GPU sharding of the extraction: one process per GPU, torch.distributed (NCCL) plumbing.

* Frames (BASELINE.json configs[3]): frames are independent units.  `frame_slice` gives every rank a contiguous block of
  frames, the rank runs them through `extract_frames` locally (meshes stay on the rank that renders them) and only the
  gradients of the parameters shared by all frames (sdf / msdf) cross the NVLink fabric: `allreduce_shared_grads`
  (dense all-reduce, or a sparse all-gather of the touched entries -- a frame touches < 1 % of the grid vertices).
* Tet ranges (configs[4]): the only O(F) stage is the classification stream.  `extract_tet_sharded` lets rank r classify
  tets [r*F/R, (r+1)*F/R) (d3h_classify_range), all-gathers the compact 32-byte valid-tet records (rank order = global
  tet order) and runs the O(surface) stages replicated on every rank (d3h_extract_from_records), so every rank holds
  the mesh of the whole grid, bit-identical to the single-GPU result.  Gradients are identical on every rank as well:
  no reduction is needed.

The reference has no multi-GPU path (single process); the parity target is the 1-GPU output.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from . import extract as E

RECORD_WORDS = _cabi.TET_RECORD_BYTES // 4   # a d3h_tet_record as int32 words


# ------------------------------------------------------------------------------------------------------------ frames
def frame_slice(n_frames: int, world: int, rank: int) -> range:
    """Contiguous block of frames owned by `rank` (sizes differ by at most one, earlier ranks take the extra frame)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def tet_range(n_tets: int, world: int, rank: int, quantum: int = 8192) -> Tuple[int, int]:
    """Tet range of `rank`: contiguous, boundaries on multiples of the compaction tile (8192 tets) so that every rank
    scans whole tiles; the last rank takes the remainder."""
    tiles = (n_tets + quantum - 1) // quantum
    lo = (tiles * rank // world) * quantum
    hi = n_tets if rank == world - 1 else (tiles * (rank + 1) // world) * quantum
    return min(lo, n_tets), min(hi, n_tets)


def allreduce_shared_grads(grads: Sequence[Optional[torch.Tensor]], group=None, sparse: bool = False,
                           sparse_threshold: float = 0.05) -> None:
    """Sum the gradients of parameters shared by all ranks' frames, in place.

    dense : one all-reduce per tensor (NCCL ring / NVLS over NVSwitch).
    sparse: every rank gathers its non-zero entries (index, value), the lists are all-gathered (padded to the longest)
            and scatter-added -- an extraction touches only the grid vertices next to the surface, so the lists are
            ~1 % of the dense size.  Falls back to dense when a rank's density exceeds `sparse_threshold`."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return
    for g in grads:
        if g is None:
            continue
        if not sparse:
            dist.all_reduce(g, group=group)
            continue
        flat = g.view(-1)
        idx = torch.nonzero(flat, as_tuple=False).view(-1)
        n_local = torch.tensor([idx.numel()], dtype=torch.int64, device=g.device)
        counts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(counts, n_local, group=group)
        counts = [int(c.item()) for c in counts]
        n_max = max(counts)
        if n_max > sparse_threshold * flat.numel():
            dist.all_reduce(g, group=group)
            continue
        if n_max == 0:
            continue
        send_i = torch.zeros(n_max, dtype=torch.int64, device=g.device)
        send_v = torch.zeros(n_max, dtype=flat.dtype, device=g.device)
        send_i[:idx.numel()] = idx
        send_v[:idx.numel()] = flat[idx]
        all_i = [torch.empty_like(send_i) for _ in range(world)]
        all_v = [torch.empty_like(send_v) for _ in range(world)]
        dist.all_gather(all_i, send_i, group=group)
        dist.all_gather(all_v, send_v, group=group)
        # every rank clears its own entries and adds ALL lists (its own included) in rank order: the same additions
        # in the same order everywhere, so replicated parameters stay bit-identical across the ranks
        flat[idx] = 0
        for r in range(world):
            if counts[r]:
                flat.index_add_(0, all_i[r][:counts[r]], all_v[r][:counts[r]])


# ------------------------------------------------------------------------------------------------------------ tet ranges
def gather_records(local: torch.Tensor, n_local: int, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather of variable-length record lists: `local` is (cap, RECORD_WORDS) int32 with `n_local` valid rows.
    Returns (records of all ranks concatenated in rank order, per-rank counts).  Works on any backend (NCCL on the GPU
    box, gloo in the CPU tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([n_local], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts)
    if n_max == 0:
        return local[:0], counts
    if local.shape[0] < n_max:   # pad the send buffer to the longest list
        send = torch.zeros((n_max, local.shape[1]), dtype=local.dtype, device=local.device)
        send[:n_local] = local[:n_local]
    else:
        send = local[:n_max].contiguous()
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send, group=group)
    return torch.cat([parts[r][:counts[r]] for r in range(world)], 0), counts


def _tet_sharded_launcher(ranges: Sequence[Tuple[int, int]], exchange: Optional[Callable], dev, group=None) -> Callable:
    """Builds the launcher that replaces d3h_extract_forward_batch in extract.forward_frames_raw.

    ranges  : tet ranges this process classifies itself (one for a real rank; several = virtual ranks on one GPU)
    exchange: (records (n,8) int32, n) -> (all records in global tet order, per-rank counts) or None (single process)
    group   : the process group of `exchange`: the overflow / total reduction must run over the same ranks"""
    L = _cabi.lib()
    c = E._FC

    def launcher(A, plan, stream):
        import ctypes as C
        assert A.shape[0] == 1, "tet-range sharding extracts one grid at a time"
        args = A[0].copy()
        cap = int(args[c["cap_valid_tets"]])
        counts_dev = torch.zeros(E._CW, dtype=torch.int64, device=dev)
        counts_pin = torch.zeros(E._CW, dtype=torch.int64).pin_memory()
        pieces, total_valid, overflow = [], 0, False
        for (lo, hi) in ranges:
            rec = torch.empty((max(cap, 1), RECORD_WORDS), dtype=torch.int32, device=dev)
            a = args.copy()
            a[c["tet_begin"]], a[c["tet_end"]] = lo, hi
            a[c["counts_host"]] = counts_pin.data_ptr()
            a[c["zero_g_pos"]:c["zero_g_msdf"] + 1] = 0
            _cabi.check(L.d3h_classify_range(a.ctypes.data, rec.data_ptr(), cap, counts_dev.data_ptr(), stream),
                        "d3h_classify_range")
            torch.cuda.current_stream(dev).synchronize()      # the one extra size read of the sharded path
            n = int(counts_pin[E._CC["n_valid_tets"]])
            total_valid += n
            if n > cap:
                overflow = True
                n = 0
            pieces.append(rec[:n])
        local = torch.cat(pieces, 0) if len(pieces) > 1 else pieces[0]
        if exchange is not None:
            # ranks that overflowed still take part in the collective (with an empty list) so that nobody hangs
            flag = torch.tensor([total_valid, int(overflow)], dtype=torch.int64, device=dev)
            import torch.distributed as dist
            dist.all_reduce(flag, group=group)
            total_valid, overflow = int(flag[0]), bool(int(flag[1]))
            merged, _ = exchange(local, local.shape[0])
        else:
            merged = local
        if overflow or total_valid > cap:
            return total_valid
        merged = merged.contiguous()
        # stage 2, replicated: n_tri / n_quad are recomputed by the library from the 4-bit codes of the records
        code = merged[:, 4]
        popc = (code & 1) + ((code >> 1) & 1) + ((code >> 2) & 1) + ((code >> 3) & 1)
        n_quad = int((popc == 2).sum())
        n_tri = merged.shape[0] - n_quad
        b = args.copy()
        _cabi.check(L.d3h_extract_from_records(b.ctypes.data, merged.data_ptr(), n_tri, n_quad, stream),
                    "d3h_extract_from_records")
        launcher.keep = (merged, counts_dev, counts_pin)   # alive until the next call / stream order
        return None

    return launcher


def extract_tet_sharded(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_negate: bool = False,
                        output_watertight_template: bool = True, group=None, virtual_ranks: Optional[int] = None):
    """One extraction whose O(F) classification is split over the ranks of `group` (or over `virtual_ranks` sequential
    sub-ranges on this GPU: the same code path without a process group, used by the single-GPU parity test).
    Returns the reference's 6-tuple, identical on every rank and bit-identical to `extract(...)`."""
    import torch.distributed as dist
    pos = E._prep_pos(pos_nx3)
    n_grid = pos.shape[0]
    sdf, msdf = E._prep_field(sdf_n, n_grid), E._prep_field(msdf_n, n_grid)
    tets = E.packed_tets(tet_fx4, n_grid)
    n_tets = tets.shape[0]
    if virtual_ranks is not None:
        ranges = [tet_range(n_tets, virtual_ranks, r) for r in range(virtual_ranks)]
        exchange = None
    else:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ranges = [tet_range(n_tets, world, rank)]
        exchange = (lambda rec, n: gather_records(rec, n, group)) if world > 1 else None
    launcher = _tet_sharded_launcher(ranges, exchange, pos.device, group)
    spec = ((((0, -1), (1, -1), (2, -1), bool(msdf_negate)),), bool(output_watertight_template), 1, launcher)
    return E._pack_result(E._ExtractFn.apply(spec, tets, pos, sdf, msdf), output_watertight_template)
