#!/usr/bin/env python
"""Device timeline of one batched forward+backward step: start/end of every kernel on every lane (CUDA events recorded
by the library around each launch; direct launches, no graph).  python profiles/lane_timeline.py [--frames 8] [--lanes 4]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import _cabi, grids, extract as E

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=128)
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--lanes", type=int, default=4)
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


for _ in range(5):
    step()
torch.cuda.synchronize()
_cabi.profile_enable(True)
step()
torch.cuda.synchronize()
tl = _cabi.profile_timeline()
_cabi.profile_enable(False)
t_end = max(b for _, b, _, _ in tl)
print(f"# {args.frames} frames on {args.lanes} lanes: {len(tl)} launches, span {t_end * 1e3:.1f} us")
print("# start_us end_us dur_us lane kernel")
for a, b, name, sid in sorted(tl):
    print(f"{a * 1e3:9.1f} {b * 1e3:9.1f} {(b - a) * 1e3:7.1f}  {sid}  {name}")
by = {}
for a, b, name, sid in tl:
    by.setdefault(name, []).append((b - a) * 1e3)
print("# mean duration per kernel under overlap (us):", {k: round(float(np.mean(v)), 1) for k, v in by.items()})
