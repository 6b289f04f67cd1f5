#!/usr/bin/env python
"""Device timeline of one batched forward inside the cached CUDA graphs (globaltimer stamps written by the kernels).
python profiles/graph_trace.py [--frames 16] [--lanes 4]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import _cabi, grids, extract as E

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--lanes", type=int, default=4)
ap.add_argument("--full", action="store_true")
ap.add_argument("--field", default="capsule")
args = ap.parse_args()
dev = torch.device("cuda:0")
pos_np, tets_np = grids.kuhn_grid(args.res)
sdf_np, msdf_np = (grids.capsule_garment_field if args.field == "capsule" else grids.sphere_plane_field)(pos_np)
N = pos_np.shape[0]
pos = torch.from_numpy(np.stack([pos_np + grids.frame_offsets(N, args.res, f) for f in range(args.frames)])).to(dev).requires_grad_(True)
sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev).requires_grad_(True)
msdf = torch.from_numpy(msdf_np).to(dev).requires_grad_(True)
tets = torch.from_numpy(tets_np).to(dev)
for _ in range(5):
    outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=args.lanes)
torch.cuda.synchronize()
_cabi.trace_enable(True)
plan = E._plan_for(dev, tets.shape[0], N)
seq0 = plan.seq
outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=args.lanes)
torch.cuda.synchronize()
tr = _cabi.trace_read()
_cabi.trace_enable(False)
rows = []
for i in range(args.frames):
    f = (seq0 + 1 + i) % 64
    for name, (a, b) in tr.get(f, {}).items():
        rows.append((a, b, name, i))
t0 = min(r[0] for r in rows)
t1 = max(r[1] for r in rows)
print(f"# {args.frames} frames on {args.lanes} lanes, graph launches: forward span {(t1 - t0) / 1e3:.1f} us = {(t1 - t0) / 1e3 / args.frames:.1f} us/frame")
order = ["prepare", "classify", "edge_scan", "edge_mark", "compact", "bucket_scan", "partition", "group_sort", "vertex_emit", "edge_emit", "poly_faces", "poly_cut", "zero"]
print("# per frame: start of prepare -> end of poly_cut (us), lane = frame % lanes; then per-kernel durations")
for i in range(args.frames):
    mine = {r[2]: r for r in rows if r[3] == i}
    if not mine:
        continue
    s = (mine["prepare"][0] - t0) / 1e3
    e = (mine["poly_cut"][1] - t0) / 1e3
    durs = " ".join(f"{k[:5]}={(mine[k][1] - mine[k][0]) / 1e3:5.1f}" for k in order if k in mine)
    print(f"frame {i:2d} lane {i % args.lanes}: {s:7.1f} -> {e:7.1f} ({e - s:6.1f})  {durs}")
by = {}
for a, b, name, i in rows:
    by.setdefault(name, []).append((b - a) / 1e3)
print("# mean duration under overlap (us):", {k: round(float(np.mean(by[k])), 1) for k in order if k in by})
if args.full:
    for a, b, name, i in sorted(rows):
        print(f"{(a - t0) / 1e3:8.1f} {(b - t0) / 1e3:8.1f} {(b - a) / 1e3:6.1f} f{i:02d} {name}")
