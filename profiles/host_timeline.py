#!/usr/bin/env python
"""Where the host time of one drop-in call goes: wraps the C-ABI entry points with perf_counter_ns stamps.
python profiles/host_timeline.py [--res 128] [--iters 200]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import _cabi, grids
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--iters", type=int, default=200)
args = ap.parse_args()
dev = torch.device("cuda:0")
pos_np, tets_np = grids.kuhn_grid(args.res)
sdf_np, msdf_np = grids.capsule_garment_field(pos_np)
pos = torch.from_numpy(pos_np).to(dev).requires_grad_(True)
sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev).requires_grad_(True)
msdf = torch.from_numpy(msdf_np).to(dev).requires_grad_(True)
tets = torch.from_numpy(tets_np).to(dev)
hm = hmSDF_Tets()
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
verts, faces, _, _, _, extra = hm(pos, sdf, msdf, tets, "cloth")
gv, gm = torch.randn_like(verts), torch.randn_like(extra["msdf"])
rows = []
for it in range(args.iters + 20):
    stamps.clear()
    pos.grad = sdf.grad = msdf.grad = None
    t0 = now()
    verts, faces, _, _, _, extra = hm(pos, sdf, msdf, tets, "cloth")
    t1 = now()
    torch.autograd.backward([verts, extra["msdf"]], [gv, gm])
    t2 = now()
    f, w, b = stamps["d3h_extract_forward_batch"][-1], stamps["d3h_wait_counts"][-1], stamps["d3h_extract_backward_batch"][-1]
    if it >= 20:
        rows.append((f[0] - t0, f[1] - f[0], w[0] - f[1], w[1] - w[0], t1 - w[1], b[0] - t1, b[1] - b[0], t2 - b[1], t2 - t0))
torch.cuda.synchronize()
names = ["py before fwd launch", "C: forward launches", "py launch->wait", "C: wait for counts", "py after wait (views, autograd)",
         "py autograd -> bwd launch", "C: backward launch", "py after bwd", "TOTAL host per fwd+bwd"]
med = np.median(np.array(rows, dtype=np.float64), axis=0) / 1e3
for n, m in zip(names, med):
    print(f"{n:36s} {m:8.1f} us")
