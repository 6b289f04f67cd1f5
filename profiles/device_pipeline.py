#!/usr/bin/env python
"""Device-only time of one fwd+bwd extraction: the kernel sequence of d3h_extract_forward + d3h_extract_backward is
captured ONCE into a CUDA graph (the library never synchronises, so it is capturable) and replayed back to back, which
removes every host gap.  This is the floor the drop-in call could reach if the host cost were zero; bench.py's `value`
includes the host.   python profiles/device_pipeline.py [--res 128] [--field capsule] [--reps 200]"""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import _cabi, grids, extract as E

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--field", default="capsule")
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--no-zero", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
pos_np, tets_np = grids.kuhn_grid(args.res)
sdf_np, msdf_np = (grids.capsule_garment_field if args.field == "capsule" else grids.sphere_plane_field)(pos_np)
pos, sdf, msdf = (torch.from_numpy(x).to(dev) for x in (pos_np, sdf_np, msdf_np))
tets = E.packed_tets(torch.from_numpy(tets_np).to(dev), pos.shape[0])
F, N = tets.shape[0], pos.shape[0]
r = E.forward_raw(pos, sdf, msdf, tets, False, True, want_grads=(True, True, True))   # learn the sizes
r = E.forward_raw(pos, sdf, msdf, tets, False, True, want_grads=(True, True, True))   # capacities settled
plan = E._plan_for(dev, F, N)
a = plan.args                                   # still holds the pointers of the last call (buffers kept alive by r)
gva = torch.randn_like(r.verts_aug); gma = torch.randn_like(r.msdf_aug)
b = _cabi.BackwardArgs()
b.pos, b.sdf, b.msdf, b.n_grid = pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr(), N
b.grads_prezeroed = 0 if args.no_zero else 1
b.tape_edges, b.tape_corners, b.tape_slots, b.tape_runs = r.tape_ptrs
b.verts_wt, b.msdf_wt = r.verts_wt.data_ptr(), r.msdf_wt.data_ptr()
b.n_verts, b.n_tri_tets, b.n_quad_tets = r.n_verts, r.n_tri, r.n_quad
b.g_verts_aug, b.g_msdf_aug = gva.data_ptr(), gma.data_ptr()
b.g_pos, b.g_sdf, b.g_msdf = (t.data_ptr() for t in r.zero_grads)
L = _cabi.lib()
if args.no_zero:
    a.zero_g_pos = a.zero_g_sdf = a.zero_g_msdf = None
s = torch.cuda.Stream()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        _cabi.check(L.d3h_extract_forward(C.byref(a), s.cuda_stream), "fwd")
        _cabi.check(L.d3h_extract_backward(C.byref(b), s.cuda_stream), "bwd")
    for _ in range(10):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(args.reps):
        g.replay()
    e1.record(s)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / args.reps * 1e3
c = r.counts
balg = grids.surface_counts_bytes(F, N, c["n_verts"], c["n_verts_aug"], c["n_faces_watertight"], c["n_faces_aug"])
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"what": "graph replay of forward+backward kernels, no host in the loop", "res": args.res, "F": F, "N": N,
                  "us_per_frame": us, "tets_per_s": F / us * 1e6, "alg_bytes": balg, "GBps": balg / us / 1e3,
                  "frac_of_measured_peak": balg / us / 1e3 / peak, "zero_in_forward": not args.no_zero}))
