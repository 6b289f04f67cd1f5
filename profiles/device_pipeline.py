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
gp, gs, gm = torch.empty_like(pos), torch.empty_like(sdf), torch.empty_like(msdf)
ptrs = [[pos.data_ptr(), sdf.data_ptr(), msdf.data_ptr()]]
gptrs = [[gp.data_ptr(), gs.data_ptr(), gm.data_ptr()]]
zero = None if args.no_zero else gptrs
for _ in range(2):   # learn the sizes, then settle the capacities
    r = E.forward_frames_raw(ptrs, [0], dev, N, tets, True, lanes=1, zero=zero, grad_ptrs=gptrs)
plan = E._plan_for(dev, F, N)
A = plan.layouts[(1, 1, 1)].A                                # still holds the pointers of the last call (buffers kept alive by r)
f0 = r.frames[0]
gva = torch.randn_like(f0.verts_aug); gma = torch.randn_like(f0.msdf_aug)
bm = r.bmat
bm[0, E._BC["g_verts_aug"]], bm[0, E._BC["g_msdf_aug"]] = gva.data_ptr(), gma.data_ptr()
if args.no_zero:
    bm[0, E._BC["msdf_negate"]] = 0             # grads_prezeroed = 0: the backward call zero-fills
L = _cabi.lib()
s = torch.cuda.Stream()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        _cabi.check(L.d3h_extract_forward_batch(A.ctypes.data, 1, 1, s.cuda_stream), "fwd")
        _cabi.check(L.d3h_extract_backward_batch(bm.ctypes.data, 1, 1, s.cuda_stream), "bwd")
    for _ in range(10):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(args.reps):
        g.replay()
    e1.record(s)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / args.reps * 1e3
c = f0.counts
balg = grids.surface_counts_bytes(F, N, c["n_verts"], c["n_verts_aug"], c["n_faces_watertight"], c["n_faces_aug"])
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"what": "graph replay of forward+backward kernels, no host in the loop", "res": args.res, "F": F, "N": N,
                  "us_per_frame": us, "tets_per_s": F / us * 1e6, "alg_bytes": balg, "GBps": balg / us / 1e3,
                  "frac_of_measured_peak": balg / us / 1e3 / peak, "zero_in_forward": not args.no_zero}))
