#!/usr/bin/env python
"""The stream kernel of the edge-scan path timed alone (d3h_profile_scan_kernel): L2-flushed and warm, 128^3 capsule frame.
Variants are chosen by environment (D3H_SCAN_ROWS, D3H_SCAN_CPW, D3H_SCAN_PHASED, D3H_SCAN_VPT): one process per variant."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from d3human_code_b200 import extract as E, single as S1  # noqa: E402
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
pos_np, sdf_np, msdf_np, tets_np = bench.make_inputs(res, "capsule")
N = pos_np.shape[0]
tets = torch.from_numpy(tets_np).to(dev)
pos = torch.from_numpy(pos_np).to(dev)
sdf = torch.from_numpy(sdf_np[:, None].copy()).to(dev)
msdf = torch.from_numpy(msdf_np).to(dev)
hm = hmSDF_Tets()
for _ in range(3):
    out = hm(pos, sdf, msdf, tets, "cloth")
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cold = [S1.profile_scan_kernel(tets, N, reps=50, flush=flush) for _ in range(3)]
warm = [S1.profile_scan_kernel(tets, N, reps=50, flush=None) for _ in range(3)]
st = E.static_edges_for(E.packed_tets(tets, N), N)
env = {k: v for k, v in os.environ.items() if k.startswith("D3H_SCAN")}
print("scan_ab", env, "V", int(out[5]["vertices_watertight"].shape[0]), "cold_us", [round(c, 2) for c in cold],
      "warm_us", [round(w, 2) for w in warm], "edges", st[2])
