#!/usr/bin/env python
"""Latency of the drop-in single call (hmSDF_Tets()(...) + backward) on one GPU: median CUDA-event time of the forward and
backward halves and wall clock per call, 128^3 capsule + garment (BASELINE configs[1]).  python profiles/single_call_timing.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import grids
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets

res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
pos, tets = grids.kuhn_grid(res)
sdf, msdf = grids.capsule_garment_field(pos)
tp = torch.tensor(pos, device=dev, requires_grad=True)
ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
tm = torch.tensor(msdf, device=dev, requires_grad=True)
tt = torch.tensor(tets, device=dev)
hm = hmSDF_Tets()
v, f, _, _, _, ex = hm(tp, ts, tm, tt, "cloth")
gv, gm = torch.randn_like(v), torch.randn_like(ex["msdf"])
fm, bm = [], []
for it in range(260):
    tp.grad = ts.grad = tm.grad = None
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    v, f, _, _, _, ex = hm(tp, ts, tm, tt, "cloth")
    b.record()
    torch.autograd.backward([v, ex["msdf"]], [gv, gm])
    c.record()
    torch.cuda.synchronize()
    if it >= 60:
        fm.append(a.elapsed_time(b)); bm.append(b.elapsed_time(c))
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(300):
    tp.grad = ts.grad = tm.grad = None
    v, f, _, _, _, ex = hm(tp, ts, tm, tt, "cloth")
    torch.autograd.backward([v, ex["msdf"]], [gv, gm])
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 300
print(f"single call res {res}: fwd {np.median(fm)*1e3:.1f} us  bwd {np.median(bm)*1e3:.1f} us  sum {1e3*(np.median(fm)+np.median(bm)):.1f} us  "
      f"back-to-back wall {wall*1e6:.1f} us/call  env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("D3H_")))
