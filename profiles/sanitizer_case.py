"""Small fwd+bwd workload for compute-sanitizer (memcheck / racecheck) on hardware: 24^3 grid, every path once:
general sort path (first call), edge-scan path (second call on), a 5-frame batch on 3 lanes, fused-frame batches with
shared topology / several rounds / mapped host positions, the fused cloth/body pair,
tet-range sharding with 3 virtual ranks, the mesh stage.  Prints 'sanitizer case ok'."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3human_code_b200 import extract as E, grids, sharding  # noqa: E402
from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets  # noqa: E402
from d3human_code_b200.render import mesh as M  # noqa: E402

dev = torch.device("cuda:0")
res = 24
pos, tets = grids.kuhn_grid(res)
sdf, msdf = grids.capsule_garment_field(pos)
tp = torch.tensor(pos, device=dev, requires_grad=True)
ts = torch.tensor(sdf[:, None], device=dev, requires_grad=True)
tm = torch.tensor(msdf, device=dev, requires_grad=True)
tt = torch.tensor(tets, device=dev)
hm = hmSDF_Tets()
for typ in ("cloth", "body", "cloth"):       # 1st: sort path, then the static table exists
    v, f, _, _, _, ex = hm(tp, ts, tm, tt, typ)
    (v.square().sum() + ex["msdf"].sum() + ex["vertices_watertight"].sum()).backward()
pb = torch.tensor(np.stack([pos + grids.frame_offsets(pos.shape[0], res, i) for i in range(5)]), device=dev, requires_grad=True)
fut = E.extract_frames_async(pb, ts, tm, tt, types=["cloth", "body", "cloth", "cloth", "body"], lanes=3)
outs = fut.result()
torch.autograd.backward([o[0].sum() + o[5]["msdf"].sum() for o in outs])
rows = E.gather_touched(pb.grad[0], fut.tape_edges(0))
pk = E.extract_frames_async(pb, ts, tm, tt, types="cloth", lanes=2).packed()
(pk.verts_aug[0, :int(pk.n_verts_aug[0])].sum()).backward()
# frames fused into one launch per kernel: shared topology (same sdf / msdf, 5 frames on 5 workspaces), then 11 frames on 4
# workspaces (three rounds), then positions read in place from pinned host memory
outs = E.extract_frames(pb, ts, tm, tt, types="cloth", lanes=5)
torch.autograd.backward([o[0].sum() + o[5]["msdf"].sum() for o in outs])
pb11 = torch.tensor(np.stack([pos + grids.frame_offsets(pos.shape[0], res, i) for i in range(11)]), device=dev, requires_grad=True)
outs = E.extract_frames(pb11, ts, tm, tt, types="body", lanes=4)
torch.autograd.backward([o[0].sum() + o[5]["msdf"].sum() for o in outs])
host = pb.detach().cpu().pin_memory()
pm = E.mapped_view(host, dev).requires_grad_(True)
outs = E.extract_frames(pm, ts, tm, tt, types="cloth", lanes=5)
torch.autograd.backward([o[0].sum() + o[5]["msdf"].sum() for o in outs])
c, b = hm.split(tp, ts, tm, tt, fused=True)
(c[0].sum() + b[0].sum()).backward()
v2, f2, *_ = sharding.extract_tet_sharded(tp, ts, tm, tt, virtual_ranks=3)
v2.sum().backward()
mesh = M.Mesh(v.detach().clone().requires_grad_(True), f)
mesh.edges
m2 = M.auto_normals(mesh)
m2.v_nrm.sum().backward()
torch.cuda.synchronize()
print("sanitizer case ok", int(v.shape[0]), int(f.shape[0]), int(rows.shape[0]))
