#!/usr/bin/env python
"""Headline benchmark: tets/s of fwd+bwd G-Shell / mSDF extraction (BASELINE.json metric) + HBM roofline fraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 128^3 Kuhn tet grid (F=12,582,912, N=2,146,689), analytic capsule-union SDF +
garment mSDF, hmSDF_Tets(type="cloth") forward + backward with upstream gradients on verts_aug and extra['msdf'].
Each rank runs `--frames-per-rank` frames per step (default 32 = two configs[3] batches); a frame is one full extraction
with its own per-frame tet-vertex offsets (seed = global frame index); the frames of a step go through `--groups`
extract_frames_async() calls of 8 frames (one autograd node and one library call per direction each, frames spread over
concurrent lanes), all launched up front so that the host work of one group overlaps the GPU work of the next.
sdf/msdf are shared, so their gradients accumulate over the rank's frames (atomics) and are all-reduced (NCCL) when
N > 1.  Weak scaling: per-GPU work is fixed.  value = N * frames_per_rank * F / (max-over-ranks device time per step).
`single_call` reports the same frames through the drop-in class one call at a time (the reference's calling pattern).

One JSON line on stdout (rank 0).  See DESIGN.md, "Measurement: keys of the bench.py JSON line".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "tets/s fwd+bwd G-Shell extraction"
UNIT = "tets/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=128, help="Kuhn grid resolution (128 = BASELINE configs[1])")
    ap.add_argument("--frames-per-rank", type=int, default=32,
                    help="frames per rank per step (BASELINE configs[3]: batches of 16 video frames; default = 2 batches)")
    ap.add_argument("--lanes", type=int, default=8, help="concurrent lanes the frames of a batch are spread over")
    ap.add_argument("--groups", type=int, default=4,
                    help="the frames of a step are issued as this many extract_frames_async batches (host / GPU pipelining)")
    ap.add_argument("--field", default="capsule", choices=["capsule", "sphere"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mesh-stage", action="store_true", help="skip the Mesh / auto_normals leg (SURVEY 8f row 2)")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the plain-PyTorch-on-this-GPU leg (SURVEY 8d)")
    ap.add_argument("--profile-steps", type=int, default=20)
    ap.add_argument("--e2e-chunk", type=int, default=2, help="frames per pipelined chunk of the end-to-end leg")
    return ap.parse_args()


def make_inputs(res, field):
    from d3human_code_b200 import grids
    pos, tets = grids.kuhn_grid(res)
    sdf, msdf = (grids.capsule_garment_field if field == "capsule" else grids.sphere_plane_field)(pos)
    return pos, sdf, msdf, tets


def workload_name(args):
    return (f"configs[1]: {args.res}^3 Kuhn grid, {'capsule-union SDF + garment mSDF' if args.field == 'capsule' else 'sphere SDF + plane mSDF'}"
            f", hmSDF_Tets(cloth) semantics fwd+bwd, {args.frames_per_rank} frame(s)/rank/step with per-frame offsets "
            f"(configs[3] batch) through extract_frames on {args.lanes} lanes")


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while a region runs."""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # pragma: no cover
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # pragma: no cover
                pass
            time.sleep(self.period)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """The reference's CPU implementation of the path, timed on this box's host cores: the numpy port in oracle/
    (the reference itself is PyTorch code that does not travel to the GPU box)."""
    if rank != 0:
        return
    from oracle import gshell_oracle as O
    cores = os.cpu_count() or 1
    pos, sdf, msdf, tets = make_inputs(args.res, args.field)
    F = tets.shape[0]
    from d3human_code_b200 import grids
    pos = pos + grids.frame_offsets(pos.shape[0], args.res, 0)

    def step():
        fwd = O.extract_forward(pos, sdf, msdf, tets, 1, True, n_threads=cores)
        gv = np.ones_like(fwd["verts_aug"])
        gm = np.ones_like(fwd["msdf"])
        O.extract_backward(fwd, gv, gm)

    # bounded sample: every step is ONE frame of the workload (fwd+bwd); cap the whole run at ~2 minutes
    step()
    t0 = time.perf_counter(); step(); one = time.perf_counter() - t0
    steps = max(1, min(args.steps, int(100.0 / max(one, 1e-3))))
    warm = max(0, min(args.warmup, 3))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    value = F / dt
    sample = f"{steps} step(s), each one full fwd+bwd extraction of one frame of the workload on the host CPU"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args), "F": int(F), "N": int(pos.shape[0])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    # D3H_BENCH_DEVICE=cpu is a TEST hook (tests/test_emu_bench.py): a control-flow dry run of this script on CPU tensors,
    # gloo and the emulated kernels of tests/emu -- it exists to catch collective mismatches between ranks without a GPU;
    # the product has no CPU path and the numbers of such a run mean nothing.
    dev_type = os.environ.get("D3H_BENCH_DEVICE", "cuda")
    if dev_type == "cuda":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU path)")
        torch.cuda.set_device(local_rank)
    dev = torch.device(dev_type, local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a mismatched collective should fail within minutes, not hold the box for the default 10
        if dev_type == "cuda":
            dist.init_process_group(backend="nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
        else:
            dist.init_process_group(backend="gloo", timeout=datetime.timedelta(seconds=120))
    from d3human_code_b200 import _cabi, grids
    from d3human_code_b200 import extract as E
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets

    pos_np, sdf_np, msdf_np, tets_np = make_inputs(args.res, args.field)
    F, N = int(tets_np.shape[0]), int(pos_np.shape[0])
    fpr = args.frames_per_rank
    frames = [rank * fpr + i for i in range(fpr)]
    hm = hmSDF_Tets()
    tets = torch.from_numpy(tets_np).to(dev)                      # int64 like hmsdf.py:207-212; packed once (static)
    sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev).requires_grad_(True)   # (N,1) like the SDF MLP output
    msdf = torch.from_numpy(msdf_np).to(dev).requires_grad_(True)
    # per-frame deformed grid vertices as ONE (B,N,3) tensor: one autograd leaf, one gradient buffer for the batch
    host_pos = torch.from_numpy(np.stack([pos_np + grids.frame_offsets(N, args.res, f) for f in frames])).pin_memory()
    host_sdf = torch.from_numpy(sdf_np[:, None].copy()).pin_memory()
    host_msdf = torch.from_numpy(msdf_np).pin_memory()
    pos = host_pos.to(dev).requires_grad_(True)
    # the frames of a step go through `--groups` extract_frames_async calls that are all launched up front: while the GPU
    # extracts group k+1 the host reads the sizes of group k, wraps its outputs and runs its backward pass
    ngroups = max(1, min(args.groups, fpr))
    gb = [(fpr * k // ngroups, fpr * (k + 1) // ngroups) for k in range(ngroups)]
    pos_groups = [pos.detach()[lo:hi].clone().requires_grad_(True) for lo, hi in gb]   # one (b,N,3) leaf per group

    # dry run: shapes of the upstream gradients (constant across steps: inputs are fixed)
    outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=args.lanes)
    counts = [dict(c) for c in E.last_counts_frames()]
    g = torch.Generator(device=dev).manual_seed(1234)
    ups_v = [torch.randn(o[0].shape, device=dev, generator=g) for o in outs]
    ups_m = [torch.randn(o[5]["msdf"].shape, device=dev, generator=g) for o in outs]
    c0 = counts[0]
    balg = sum(grids.surface_counts_bytes(F, N, c["n_verts"], c["n_verts_aug"], c["n_faces_watertight"], c["n_faces_aug"])
               for c in counts) / len(counts)
    del outs

    def local_step():
        """This rank's share of a step, no collective: every group of frames is one extract_frames_async call (one
        library call, concurrent lanes) and one backward call; gradients of the shared sdf / msdf are summed over the
        frames by the kernels and over the groups by autograd."""
        sdf.grad = msdf.grad = None
        for pg in pos_groups:
            pg.grad = None
        futs = [E.extract_frames_async(pg, sdf, msdf, tets, types="cloth", lanes=args.lanes) for pg in pos_groups]
        for fut, (lo, hi) in zip(futs, gb):
            outs = fut.result()
            torch.autograd.backward([o[0] for o in outs] + [o[5]["msdf"] for o in outs], ups_v[lo:hi] + ups_m[lo:hi])

    def step():
        """One training-step's worth of extraction: local_step() + the sum of the shared gradients over the ranks (NCCL).
        Every rank must call it the same number of times."""
        local_step()
        if world > 1:
            dist.all_reduce(sdf.grad)
            dist.all_reduce(msdf.grad)

    pos_single = [pos.detach()[i].clone().requires_grad_(True) for i in range(min(fpr, 4))]   # separate (N,3) leaves

    def single_call_step():
        """The reference's own calling pattern: one drop-in call + backward per frame (hmsdf.py:548)."""
        sdf.grad = msdf.grad = None
        for p, gv, gm in zip(pos_single, ups_v, ups_m):
            p.grad = None
            verts, faces, _, _, v_tng, extra = hm(p, sdf, msdf, tets, "cloth")
            torch.autograd.backward([verts, extra["msdf"]], [gv, gm])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        tmax = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()) / steps

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = E.launch_counter()
    ms_step = timed(step, args.steps)
    launches_timed = E.launch_counter() - launches0      # library kernels enqueued by this rank in the timed region
    value = world * fpr * F / (ms_step * 1e-3)

    # ---- the same frames through the drop-in class, one call + backward per frame (no batching, no lanes) ----
    single = None
    if world == 1:
        k_single = max(3, min(args.steps, 50))
        for _ in range(3):
            single_call_step()
        ms_single = timed(single_call_step, k_single) / len(pos_single)
        single = {"ms_per_frame": ms_single, "value": F / (ms_single * 1e-3), "unit": UNIT, "steps": k_single,
                  "note": "hmSDF_Tets()(...) + backward per frame, strictly serial on the host (the reference's calling pattern)"}
        # forward and backward separately: CUDA events around each half of one drop-in call, median
        try:
            f_ms, b_ms = [], []
            p1, gv1, gm1 = pos_single[0], ups_v[0], ups_m[0]
            for _ in range(min(max(k_single, 10), 100)):
                sdf.grad = msdf.grad = p1.grad = None
                ea, eb_, ec = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                ea.record()
                verts, faces, _, _, v_tng, extra = hm(p1, sdf, msdf, tets, "cloth")
                eb_.record()
                torch.autograd.backward([verts, extra["msdf"]], [gv1, gm1])
                ec.record()
                barrier()
                f_ms.append(ea.elapsed_time(eb_))
                b_ms.append(eb_.elapsed_time(ec))
            single["fwd_ms_median"], single["bwd_ms_median"] = float(np.median(f_ms)), float(np.median(b_ms))
        except Exception as exc:  # noqa: BLE001
            single["fwd_bwd_split_error"] = f"{type(exc).__name__}: {exc}"[:200]

    # ---- per-kernel device time (CUDA events on the launching stream, recorded by the library; one frame at a time so
    # that every kernel is timed alone) ----
    prof = {}
    if rank == 0:
        _cabi.profile_enable(True)
        nprof = max(1, min(args.profile_steps, 128 // len(pos_single)))
        for _ in range(nprof):
            single_call_step()
        torch.cuda.synchronize()
        prof = _cabi.profile_read()
        _cabi.profile_enable(False)
    # ---- device-side stamps (globaltimer written by the kernels themselves) of one batched step inside the cached
    # graphs: duration of every kernel as it runs in the timed region, overlapped with the other lanes ----
    dev_trace = None
    if rank == 0:
        try:
            _cabi.trace_enable(True)
            plan = E._plan_for(pos.device, F, N)
            seq0 = plan.seq
            local_step()          # rank 0 only: must not contain a collective
            torch.cuda.synchronize()
            tr = _cabi.trace_read()
            _cabi.trace_enable(False)
            durs = {}
            t_lo, t_hi = None, None
            for i in range(fpr):
                for name, (a, b) in tr.get((seq0 + 1 + i) % 64, {}).items():
                    durs.setdefault(name, []).append((b - a) / 1e3)
                    t_lo = a if t_lo is None else min(t_lo, a)
                    t_hi = b if t_hi is None else max(t_hi, b)
            if durs:
                dev_trace = {"us_mean": {k: round(float(np.mean(v)), 2) for k, v in durs.items()},
                             "forward_span_us": round((t_hi - t_lo) / 1e3, 1), "frames": fpr}
        except Exception as exc:  # pragma: no cover
            dev_trace = {"error": str(exc)}
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- end to end through the public API with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        # The batch is cut into chunks of `--e2e-chunk` frames that flow through three streams: H2D copies of chunk
        # k+1 and D2H copies of chunk k-1 run under the extraction of chunk k (PCIe is full duplex, ~55 GB/s each way).
        chunk = max(1, min(args.e2e_chunk, fpr))
        bounds = [(i, min(i + chunk, fpr)) for i in range(0, fpr, chunk)]
        d_sdf = torch.empty_like(sdf)
        d_msdf = torch.empty_like(msdf)
        d_pos = [torch.empty_like(pos[lo:hi]) for lo, hi in bounds]
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        outs_host = None
        k_e2e = max(3, min(args.steps, 20))

        def e2e_step():
            nonlocal outs_host
            h2d = d2h = 0
            cur = torch.cuda.current_stream()
            s_in.wait_stream(cur)
            ev_in = []
            with torch.cuda.stream(s_in):
                for dst, src in ((d_sdf, host_sdf), (d_msdf, host_msdf)):
                    dst.requires_grad_(False)
                    dst.copy_(src, non_blocking=True)
                    dst.requires_grad_(True)
                    dst.grad = None
                h2d += host_sdf.numel() * 4 + host_msdf.numel() * 4
                for dp, (lo, hi) in zip(d_pos, bounds):
                    dp.requires_grad_(False)
                    dp.copy_(host_pos[lo:hi], non_blocking=True)
                    dp.requires_grad_(True)
                    dp.grad = None
                    h2d += dp.numel() * 4
                    ev = torch.cuda.Event()
                    ev.record(s_in)
                    ev_in.append(ev)
            keep, res_all = [], []
            for k, (dp, (lo, hi)) in enumerate(zip(d_pos, bounds)):
                cur.wait_event(ev_in[k])
                outs = E.extract_frames(dp, d_sdf, d_msdf, tets, types="cloth", lanes=args.lanes)
                torch.autograd.backward([o[0] for o in outs] + [o[5]["msdf"] for o in outs],
                                        ups_v[lo:hi] + ups_m[lo:hi])
                res = []
                for o in outs:
                    res += [o[0].detach(), o[1], o[5]["msdf"].detach()]
                res.append(dp.grad)
                if k == len(bounds) - 1:
                    res += [d_sdf.grad, d_msdf.grad]
                s_out.wait_stream(cur)
                res_all.append(res)
                if outs_host is not None:
                    with torch.cuda.stream(s_out):
                        for h, t in zip(outs_host[k], res):
                            h.copy_(t, non_blocking=True)
                            d2h += t.numel() * t.element_size()
                keep.append(outs)
            if outs_host is None:   # first call: allocate the pinned result buffers, copy without overlap
                outs_host = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in res] for res in res_all]
                for hs, res in zip(outs_host, res_all):
                    for h, t in zip(hs, res):
                        h.copy_(t, non_blocking=True)
                        d2h += t.numel() * t.element_size()
            torch.cuda.synchronize()
            return h2d, d2h

        for _ in range(3):
            h2d_b, d2h_b = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / k_e2e], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * fpr * F / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d_b),
               "d2h_bytes_per_step": int(d2h_b), "steps": k_e2e, "ms_per_step": float(dt.item()) * 1e3,
               "chunk_frames": chunk,
               "note": "extract_frames() with inputs copied from pinned host memory each step (pos per frame, sdf, "
                       "msdf); per frame verts_aug, faces_aug, msdf and the dense pos gradient, plus the sdf / msdf "
                       "gradients, copied back to pinned host memory; static tet indices stay resident; chunks of "
                       "frames pipelined over H2D / compute / D2H streams"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (classify: the only O(F) kernel) ----
    peak, peak_src = FALLBACK_HBM_GBS, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        pass
    roofline = None
    kern = {}
    if prof:
        kern = {k: {"ms_total": round(v[0], 4), "launches": v[1], "us_avg": round(1e3 * v[0] / v[1], 3)} for k, v in prof.items()}
        dom = "edge_scan" if "edge_scan" in prof else "classify"
        ms, n = prof.get(dom, (0.0, 0))
        dom_bytes = 16.0 * F
        if dom == "edge_scan":   # 4 B per edge (larger endpoint) + 4 B per vertex (CSR offsets) + the sign bitmap
            st = E.static_edges_for(E.packed_tets(tets, N), N)
            dom_bytes = 4.0 * st[2] + 4.0 * (N + 1) + N / 8.0
        if n and ms > 0:
            t = ms / n * 1e-3
            achieved = dom_bytes / t / 1e9
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "classify_traffic.json")) as fh:
                    tj = json.load(fh)
                    if int(tj.get("F", 0)) == F and tj.get("kernel", "classify") == dom:
                        traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
            roofline = {"kernel": dom + "_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": int(dom_bytes), "us_per_launch": t * 1e6,
                        "timing": "CUDA events recorded around each launch on its stream, one frame at a time (includes "
                                  "the ~3-5 us event / launch gap)"}
            if dev_trace and "us_mean" in dev_trace and dom in dev_trace["us_mean"]:
                td = dev_trace["us_mean"][dom] * 1e-6
                roofline["device_timer"] = {"us_per_launch": td * 1e6, "achieved": dom_bytes / td / 1e9,
                                            "frac": dom_bytes / td / 1e9 / peak,
                                            "note": "first block start to last block exit (%globaltimer), inside the "
                                                    "batched step with the other lanes running"}
    dev_ms_frame = sum(v[0] for v in prof.values()) / max(nprof * len(pos_single), 1) if prof else None
    path_roofline = {"algorithmic_bytes_per_frame": int(balg),
                     "achieved_GBps_step": balg * fpr / (ms_step * 1e-3) / 1e9,
                     "frac_step": balg * fpr / (ms_step * 1e-3) / 1e9 / peak,
                     "kernel_ms_per_frame": dev_ms_frame,
                     "frac_kernels_only": (balg / (dev_ms_frame * 1e-3) / 1e9 / peak) if dev_ms_frame else None}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import gshell_oracle as O
        cores = os.cpu_count() or 1
        p0 = host_pos[0].numpy()

        def cpu_step():
            fwd = O.extract_forward(p0, sdf_np, msdf_np, tets_np, 1, True, n_threads=cores)
            O.extract_backward(fwd, np.ones_like(fwd["verts_aug"]), np.ones_like(fwd["msdf"]))

        cpu_step()
        reps, best, t_all = 0, 1e9, time.perf_counter()
        while reps < 20 and time.perf_counter() - t_all < 15.0:
            t0 = time.perf_counter(); cpu_step(); best = min(best, time.perf_counter() - t0); reps += 1
        cpu_baseline = {"value": F / best, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"best of {reps} full fwd+bwd extractions of one frame of the same workload, numpy oracle "
                                  f"with the O(F) stage on {cores} threads",
                        "ms_per_frame": best * 1e3}
        try:  # the same path as plain PyTorch ops on the host cores: how the reference itself behaves on a CPU
            from oracle import gshell_torch as GT
            cp = torch.from_numpy(p0.copy()).requires_grad_(True)
            cs, cm = torch.from_numpy(sdf_np[:, None].copy()).requires_grad_(True), torch.from_numpy(msdf_np.copy()).requires_grad_(True)
            ct = torch.from_numpy(tets_np)
            tbest = 1e9
            for _ in range(2):
                cp.grad = cs.grad = cm.grad = None
                t0 = time.perf_counter()
                v, _f, _, _, _, ex = GT.extract(cp, cs, cm, ct, 1, True)
                torch.autograd.backward([v, ex["msdf"]], [torch.ones_like(v), torch.ones_like(ex["msdf"])])
                tbest = min(tbest, time.perf_counter() - t0)
            cpu_baseline["torch_port"] = {"value": F / tbest, "unit": UNIT, "ms_per_frame": tbest * 1e3,
                                          "threads": torch.get_num_threads(),
                                          "note": "oracle/gshell_torch.py on CPU tensors, best of 2: the reference's op "
                                                  "sequence on the host (the numpy figure above is the faster, conservative one)"}
        except Exception as exc:  # noqa: BLE001
            cpu_baseline["torch_port"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    # ---- the stage right behind the extraction: Mesh.edges + auto_normals fwd / bwd of this package on
    # frame 0's surfaces against the same PyTorch ops on the device.  Last leg, never fatal: the headline is complete.
    mesh_stage = None
    if world == 1 and not args.no_mesh_stage:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_mesh_bench", os.path.join(ROOT, "profiles", "mesh_bench.py"))
            mb = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mb)
            with torch.no_grad():
                verts, faces, _, _, _, extra = hm(pos_single[0].detach(), sdf.detach(), msdf.detach(), tets, "cloth")
            l0 = mb.mesh.launch_counter()
            mesh_stage = mb.measure([("open", verts.detach(), faces),
                                     ("watertight", extra["vertices_watertight"].detach(), extra["faces_watertight"])],
                                    max(3, min(args.steps, 100)), sync=None if dev_type == "cuda" else "cpu")
            mesh_stage["gpu_launches"] = mb.mesh.launch_counter() - l0
            mesh_stage["unit"] = "us per call (median, CUDA events; edges include the host's size read)"
        except Exception as exc:  # noqa: BLE001
            mesh_stage = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- the reference's way on the same GPU: the path as plain PyTorch ops + autograd (oracle/gshell_torch.py, a port
    # pinned against the reference's golden vectors; the reference tree itself is not on this box).  SURVEY 8(d).  Never fatal.
    torch_baseline = None
    if world == 1 and not args.no_torch_baseline:
        try:
            from oracle import gshell_torch as GT
            tp = pos_single[0].detach().clone().requires_grad_(True)
            ts_, tm_ = sdf.detach().clone().requires_grad_(True), msdf.detach().clone().requires_grad_(True)

            def torch_step():
                tp.grad = ts_.grad = tm_.grad = None
                v, f_, _, _, _, ex = GT.extract(tp, ts_, tm_, tets, 1, True)
                torch.autograd.backward([v, ex["msdf"]], [ups_v[0], ups_m[0]])

            for _ in range(2):
                torch_step()
            reps, t_all = 0, time.perf_counter()
            barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            while reps < 20 and time.perf_counter() - t_all < 10.0:
                torch_step()
                reps += 1
            ev1.record()
            barrier()
            ms = ev0.elapsed_time(ev1) / max(reps, 1)
            torch_baseline = {"value": F / (ms * 1e-3), "unit": UNIT, "ms_per_frame": ms, "kind": "port", "reps": reps,
                              "note": "the same extraction fwd+bwd of one frame as plain PyTorch ops + autograd on this GPU "
                                      "(boolean-mask compactions, torch.unique(dim=0), scatter_add, the full UV table), "
                                      "the way the reference runs it; oracle/gshell_torch.py"}
        except Exception as exc:  # noqa: BLE001
            torch_baseline = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "F": F, "N": N, "frames_per_step": world * fpr,
                       "counts_frame0": c0, "l2": "tet index stream is 16*F = %d MB > 126 MB L2; no explicit flush" % (16 * F // 1000000),
                       "lanes": args.lanes, "groups": ngroups, "parallelism": f"frames x{world}" if world > 1 else "single GPU"},
            "roofline": roofline, "path_roofline": path_roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "device_trace": dev_trace, "single_call": single, "gpu_launches": int(launches_timed), "kernels": kern,
            "mesh_stage": mesh_stage, "torch_gpu_baseline": torch_baseline, "clocks": sampler.result()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
