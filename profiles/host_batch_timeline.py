#!/usr/bin/env python
"""Where the host time of one batched step (extract_frames + backward) goes.
python profiles/host_batch_timeline.py [--res 128] [--frames 16] [--lanes 4] [--iters 50]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import _cabi, grids, extract as E

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--lanes", type=int, default=4)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
dev = torch.device("cuda:0")
pos_np, tets_np = grids.kuhn_grid(args.res)
sdf_np, msdf_np = grids.capsule_garment_field(pos_np)
N = pos_np.shape[0]
pos = torch.from_numpy(np.stack([pos_np + grids.frame_offsets(N, args.res, f) for f in range(args.frames)])).to(dev).requires_grad_(True)
sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev).requires_grad_(True)
msdf = torch.from_numpy(msdf_np).to(dev).requires_grad_(True)
tets = torch.from_numpy(tets_np).to(dev)
L = _cabi.lib()
stamps = {}
now = time.perf_counter_ns


class Wrap:
    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, *a):
        t0 = now(); r = self.fn(*a); t1 = now()
        stamps.setdefault(self.name, []).append((t0, t1))
        return r


for name in ("d3h_extract_forward_batch", "d3h_wait_counts", "d3h_extract_backward_batch"):
    setattr(L, name, Wrap(name, getattr(L, name)))
outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=args.lanes)
gv = [torch.randn_like(o[0]) for o in outs]
gm = [torch.randn_like(o[5]["msdf"]) for o in outs]
rows = []
for it in range(args.iters + 10):
    stamps.clear()
    sdf.grad = msdf.grad = pos.grad = None
    t0 = now()
    outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=args.lanes)
    t1 = now()
    torch.autograd.backward([o[0] for o in outs] + [o[5]["msdf"] for o in outs], gv + gm)
    t2 = now()
    f = stamps["d3h_extract_forward_batch"][-1]
    w = stamps["d3h_wait_counts"]
    b = stamps["d3h_extract_backward_batch"][-1]
    wait_total = sum(x[1] - x[0] for x in w)
    if it >= 10:
        rows.append((f[0] - t0, f[1] - f[0], w[0][0] - f[1], wait_total, (w[-1][1] - w[0][0]) - wait_total, t1 - w[-1][1],
                     b[0] - t1, b[1] - b[0], t2 - b[1], t2 - t0))
torch.cuda.synchronize()
names = ["py before fwd launch", "C: forward batch launch", "py launch->first wait", "C: waits for counts (sum)",
         "py between waits", "py after waits (views, autograd)", "py autograd -> bwd launch", "C: backward batch launch",
         "py after bwd", "TOTAL host per step"]
med = np.median(np.array(rows, dtype=np.float64), axis=0) / 1e3
for n, m in zip(names, med):
    print(f"{n:36s} {m:8.1f} us   ({m / args.frames:6.1f} us/frame)")
