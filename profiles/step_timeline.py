#!/usr/bin/env python
"""Kernel / copy timeline of one bench step (32 frames in groups, packed API) from the torch profiler (CUPTI):
when does every kernel start and end, which stream, and how long does the host take to issue the step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from d3human_code_b200 import extract as E, grids  # noqa: E402

# (NGROUPS, not GROUPS: bash keeps a read-only array of that name)
frames, groups, lanes = int(os.environ.get("FRAMES", 32)), int(os.environ.get("NGROUPS", 2)), 8
dev = torch.device("cuda:0")
pos_np, sdf_np, msdf_np, tets_np = bench.make_inputs(128, "capsule")
N = pos_np.shape[0]
tets = torch.from_numpy(tets_np).to(dev)
sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev).requires_grad_(True)
msdf = torch.from_numpy(msdf_np).to(dev).requires_grad_(True)
flat = torch.zeros(2 * N, device=dev)
pos = torch.from_numpy(np.stack([pos_np + grids.frame_offsets(N, 128, f) for f in range(frames)])).to(dev)
gb = bench.group_bounds(frames, groups)
pgs = [pos[lo:hi].clone().requires_grad_(True) for lo, hi in gb]
outs = E.extract_frames(pos, sdf, msdf, tets, types="cloth", lanes=lanes)
va = max(o[0].shape[0] for o in outs)
pad = 2 * va + 4096
uv = [torch.randn((hi - lo, pad, 3), device=dev) for lo, hi in gb]
um = [torch.randn((hi - lo, pad), device=dev) for lo, hi in gb]
del outs


def step():
    flat.zero_()
    sdf.grad, msdf.grad = flat[:N].view(N, 1), flat[N:]
    for pg in pgs:
        pg.grad = None
    futs = [E.extract_frames_async(pg, sdf, msdf, tets, types="cloth", lanes=lanes) for pg in pgs]
    for k, fut in enumerate(futs):
        pk = fut.packed()
        c = pk.verts_aug.shape[1]
        torch.autograd.backward([pk.verts_aug, pk.msdf], [uv[k][:, :c], um[k][:, :c]])


for _ in range(10):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
t_issue = (time.perf_counter() - t0) / 50
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 50
print(f"step: host issue {t_issue * 1e3:.3f} ms, with the device {t_all * 1e3:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t_first = ev[0].time_range.start
last = len(ev) // 3
sel = ev[last:2 * last]         # the middle step, roughly
base = sel[0].time_range.start
busy = 0.0
for e in sel:
    busy += e.time_range.end - e.time_range.start
    print(f"{e.time_range.start - base:9.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:90]}")
print(f"kernels of the step: {len(sel)}, busy {busy:.1f} us, span {sel[-1].time_range.end - base:.1f} us")
