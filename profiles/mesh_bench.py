"""Mesh stage on one GPU: `Mesh(...).edges` and `auto_normals` forward / backward of this package against
the same operations written with plain PyTorch ops on the same device (what the reference's render/mesh.py:240-250 and
:418-441 execute), on the surfaces of the 128^3 capsule / garment extraction (BASELINE.json configs[1] inputs).

    python profiles/mesh_bench.py [--res 128] [--reps 200]

Prints one JSON line.  Times are CUDA-event times per call in microseconds (median of `reps`, after 10 warm-ups); the
edge timings include the host's size read.  Not part of bench.py's headline; kept next to it for the row's measurement.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3human_code_b200 import _cabi, grids  # noqa: E402
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets  # noqa: E402
from d3human_code_b200.render import mesh  # noqa: E402


def torch_edges(f):  # the op sequence of render/mesh.py:240-250
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    return torch.unique(torch.sort(e, dim=1).values, dim=0)


def torch_normals(p, f):  # the op sequence of render/mesh.py:420-441
    i0, i1, i2 = f[:, 0], f[:, 1], f[:, 2]
    fn = torch.linalg.cross(p[i1] - p[i0], p[i2] - p[i0])
    n = torch.zeros_like(p)
    for i in (i0, i1, i2):
        n.scatter_add_(0, i[:, None].repeat(1, 3), fn)
    d = (n * n).sum(-1, keepdim=True)
    n = torch.where(d > 1e-20, n, torch.tensor([0.0, 0.0, 1.0], device=p.device))
    return n / torch.sqrt(torch.clamp((n * n).sum(-1, keepdim=True), min=1e-20))


def timed(fn, reps, sync=None):
    """median CUDA-event time of one call in microseconds (wall clock when there is no CUDA device: bench dry runs)"""
    cuda = torch.cuda.is_available() and sync is None
    for _ in range(min(10, reps)):
        fn()
    if cuda:
        torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        else:
            import time
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e6)
    return float(np.median(ts))


def measure(surfaces, reps, sync=None):
    """surfaces: [(name, v (V,3) detached, f (F,3))] on the device -> {name: sizes and per-call times in microseconds}.
    Also checks the results against the torch ops (edges equal, normals to 1e-3: ill-conditioned vertices differ)."""
    keep = mesh.share_normals
    mesh.share_normals = False      # time the kernels, not the memo of repeated calls
    out = {}
    try:
        for name, v, f in surfaces:
            g = torch.randn_like(v)
            row = {"V": int(v.shape[0]), "F": int(f.shape[0])}

            def ours_edges():
                mesh._edge_cache.clear()     # time the kernels, not the per-faces-tensor cache
                return mesh.Mesh(v, f).edges

            row["E"] = int(ours_edges().shape[0])
            row["edges_equal_torch"] = bool(torch.equal(ours_edges(), torch_edges(f)))
            row["edges_us"] = timed(ours_edges, reps, sync)
            row["edges_torch_us"] = timed(lambda: torch_edges(f), reps, sync)
            row["normals_fwd_us"] = timed(lambda: mesh.vertex_normals(v, f), reps, sync)
            row["normals_fwd_torch_us"] = timed(lambda: torch_normals(v, f), reps, sync)
            row["normals_median_abs_diff_vs_torch"] = float((mesh.vertex_normals(v, f) - torch_normals(v, f)).abs().median())

            def fb(fn):
                p = v.clone().requires_grad_(True)
                (fn(p, f) * g).sum().backward()

            row["normals_fwd_bwd_us"] = timed(lambda: fb(mesh.vertex_normals), reps, sync)
            row["normals_fwd_bwd_torch_us"] = timed(lambda: fb(torch_normals), reps, sync)
            # device time of each entry point alone (events recorded by the library around its launches)
            _cabi.profile_read()
            _cabi.profile_enable(True)
            for _ in range(5):
                ours_edges()
                fb(mesh.vertex_normals)
            _cabi.profile_enable(False)
            row["device_us"] = {k: 1e3 * ms / max(n, 1) for k, (ms, n) in _cabi.profile_read().items() if k.startswith("mesh_")}
            out[name] = row
    finally:
        mesh.share_normals = keep
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=128)
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pos, tets = grids.kuhn_grid(args.res)
    sdf, msdf = grids.capsule_garment_field(pos)
    verts, faces, _, _, _, extra = hmSDF_Tets()(torch.tensor(pos, device=dev), torch.tensor(sdf, device=dev),
                                                torch.tensor(msdf, device=dev), torch.tensor(tets, device=dev), "cloth")
    l0 = mesh.launch_counter()
    out = {"workload": f"kuhn{args.res}_capsule_garment surfaces", "reps": args.reps}
    out.update(measure([("open", verts.detach(), faces),
                        ("watertight", extra["vertices_watertight"].detach(), extra["faces_watertight"])], args.reps))
    out["gpu_launches"] = mesh.launch_counter() - l0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
