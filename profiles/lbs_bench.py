#!/usr/bin/env python
"""Skinning stage on one GPU at the sizes of D3-Human: P = 230 286 rows of verts_aug (128^3 capsule /
garment extraction: 82 % exact zeros), SMPL-X-sized rig (10 475 template vertices, 55 joints), forward and forward +
backward, this package against the op sequence of deform/smplx_exavatar_deformer.py:363-421, 472-476 in plain PyTorch on
the same device (brute-force nearest vertex in chunks instead of pytorch3d's knn_points).  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.test_lbs_oracle import synthetic_rig


def torch_lbs(pts, template, w, init_a, a, trans):
    p = pts.reshape(1, -1, 3)
    idx = torch.cat([torch.cdist(p[0, i:i + 8192], template).argmin(1) for i in range(0, p.shape[1], 8192)])     # knn_points, K = 1
    w_pts = w[idx][None]                                                            # (1,P,J)   (:371-381)

    def apply(x, mats, inverse):                                                    # apply_lbs_inverse :398-421
        ph = torch.cat([x, torch.ones_like(x[..., :1])], 2)
        m = (mats[None, None] * w_pts[..., None, None]).sum(2)                      # (1,P,4,4)
        if inverse:
            m = torch.inverse(m)
        return torch.matmul(m, ph.unsqueeze(-1))[..., :3, 0]
    can = apply(p, init_a, True)
    return (apply(can, a, False) + trans.reshape(1, 1, 3)).reshape(-1, 3)


def measure(reps=5, dev=None):
    from d3human_code_b200 import grids
    from d3human_code_b200.deform.lbs import LinearBlendSkinning
    from d3human_code_b200.geometry.hmsdf_tets_split import hmSDF_Tets
    dev = dev or torch.device("cuda:0")
    pos, tets = grids.kuhn_grid(128)
    sdf, msdf = grids.capsule_garment_field(pos)
    with torch.no_grad():
        verts = hmSDF_Tets()(torch.tensor(pos, device=dev), torch.tensor(sdf, device=dev), torch.tensor(msdf, device=dev),
                             torch.tensor(tets, device=dev), "cloth")[0]
    template, w, init_a, a, trans = synthetic_rig(1, 10475, 55)
    template = (template * 0.9).astype(np.float32)
    T, W, IA, A, TR = (torch.tensor(t, device=dev) for t in (template, w, init_a, a, trans))
    rig = LinearBlendSkinning(T, W, IA)
    pts = verts.detach().clone().requires_grad_(True)
    A.requires_grad_(True)
    g = torch.randn_like(pts)

    def timed(fn, backward):
        def run():
            pts.grad = A.grad = None
            out = fn()
            if backward:
                (out * g).sum().backward()
            return out
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    ours = lambda: rig.lbs_transform(pts, A, TR)
    ref = lambda: torch_lbs(pts, T, W, IA, A, TR)
    with torch.no_grad():
        diff = float((ours() - ref()).abs().max())
    l0 = LinearBlendSkinning.launches
    res = {"points": int(pts.shape[0]), "zero_rows": float((verts.abs().sum(1) == 0).float().mean()), "template": 10475, "joints": 55,
           "max_abs_diff_vs_torch": diff, "unit": "ms per call (median, CUDA events)"}
    with torch.no_grad():
        res["ours_fwd_ms"], res["torch_fwd_ms"] = timed(ours, False), timed(ref, False)
    res["ours_fwd_bwd_ms"], res["torch_fwd_bwd_ms"] = timed(ours, True), timed(ref, True)
    res["gpu_launches"] = LinearBlendSkinning.launches - l0
    return res


if __name__ == "__main__":
    print(json.dumps(measure()))
