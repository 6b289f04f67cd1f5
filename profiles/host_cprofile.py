#!/usr/bin/env python
"""cProfile of the batched step (host side): python profiles/host_cprofile.py [--frames 16] [--lanes 4]"""
import argparse, cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import grids, extract as E

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--lanes", type=int, default=4)
ap.add_argument("--iters", type=int, default=100)
args = ap.parse_args()
dev = torch.device("cuda:0")
pos_np, tets_np = grids.kuhn_grid(args.res)
sdf_np, msdf_np = grids.capsule_garment_field(pos_np)
N = pos_np.shape[0]
pos = torch.from_numpy(np.stack([pos_np + grids.frame_offsets(N, args.res, f) for f in range(args.frames)])).to(dev).requires_grad_(True)
sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev).requires_grad_(True)
msdf = torch.from_numpy(msdf_np).to(dev).requires_grad_(True)
tets = torch.from_numpy(tets_np).to(dev)
outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=args.lanes)
gv = [torch.randn_like(o[0]) for o in outs]
gm = [torch.randn_like(o[5]["msdf"]) for o in outs]


def step():
    sdf.grad = msdf.grad = pos.grad = None
    outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=args.lanes)
    torch.autograd.backward([o[0] for o in outs] + [o[5]["msdf"] for o in outs], gv + gm)


for _ in range(10):
    step()
pr = cProfile.Profile()
pr.enable()
for _ in range(args.iters):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
