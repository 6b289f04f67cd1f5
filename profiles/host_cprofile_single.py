#!/usr/bin/env python
"""cProfile of the drop-in single call + backward (host side)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import grids
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
dev = torch.device("cuda:0")
pos_np, tets_np = grids.kuhn_grid(128)
sdf_np, msdf_np = grids.capsule_garment_field(pos_np)
pos = torch.from_numpy(pos_np).to(dev).requires_grad_(True)
sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev).requires_grad_(True)
msdf = torch.from_numpy(msdf_np).to(dev).requires_grad_(True)
tets = torch.from_numpy(tets_np).to(dev)
hm = hmSDF_Tets()
verts, faces, _, _, _, extra = hm(pos, sdf, msdf, tets, "cloth")
gv, gm = torch.randn_like(verts), torch.randn_like(extra["msdf"])


def step():
    pos.grad = sdf.grad = msdf.grad = None
    verts, faces, _, _, _, extra = hm(pos, sdf, msdf, tets, "cloth")
    torch.autograd.backward([verts, extra["msdf"]], [gv, gm])


for _ in range(20):
    step()
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
