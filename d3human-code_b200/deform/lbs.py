"""Linear-blend skinning of the extracted vertices: the arithmetic of `SMPLX_Deformer.interpolate_weights`,
`apply_lbs_inverse`, `lbs_forward_inverse` and of the last three lines of `lbs_forward`
(deform/smplx_exavatar_deformer.py:363-430, 472-476) behind the C ABI of include/d3h_lbs.h.  SURVEY.md section 8(f) row 4.

The reference class itself loads the SMPL-X model files and runs the SMPL-X layer (`self.layer.forward`, :456-470) to get
the joint transforms `A` of the frame; that part stays in the reference.  A maintainer keeps `SMPLX_Deformer`, builds one
`LinearBlendSkinning(self.vs_template, self.lbs_weights, self.init_A, self.k)` at the end of `initialize` (:235) and
replaces :472-476 by `return self._lbs.lbs_transform(pts, A, trans)`.

What changes: the reference blends J 4x4 matrices, inverts a 4x4 and multiplies PER POINT, twice, on every row of
verts_aug (>= 80 % of them exact zeros: the unreferenced boundary slots).  With `k = 1` (the class hard-codes it, :39) the
blended matrix only depends on the point's nearest template vertex: the tables are built per template vertex (10 475
rows), the zero rows share one nearest-vertex search, and a point costs two 4x4 products.  Gradients: to the points, to
`A` and to `trans` (the index carries none; `init_A` is a constant of the run).  CUDA tensors only.
"""
from __future__ import annotations

import torch

from .. import _cabi


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class _TransformFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, a, trans, idx, b_inv, lbs_weights):
        L = _cabi.lib()
        dev = pts.device
        p, vt, nj = pts.shape[0], lbs_weights.shape[0], lbs_weights.shape[1]
        f32 = torch.float32
        with torch.cuda.device(dev):
            b_pose = torch.empty((vt, 16), dtype=f32, device=dev)
            _cabi.check(L.d3h_lbs_blend(lbs_weights.data_ptr(), a.data_ptr(), vt, nj, 0, b_pose.data_ptr(), _st(dev)), "d3h_lbs_blend")
            can = torch.empty((p, 3), dtype=f32, device=dev)
            out = torch.empty((p, 3), dtype=f32, device=dev)
            _cabi.check(L.d3h_lbs_apply(pts.data_ptr(), p, idx.data_ptr(), b_inv.data_ptr(), b_pose.data_ptr(),
                                        trans.data_ptr() if trans is not None else None, can.data_ptr(), out.data_ptr(), _st(dev)),
                        "d3h_lbs_apply")
        ctx.save_for_backward(pts, idx, b_inv, b_pose, can, lbs_weights)
        ctx.has_trans = trans is not None
        LinearBlendSkinning.launches += 2
        return out

    @staticmethod
    def backward(ctx, g):
        pts, idx, b_inv, b_pose, can, lbs_weights = ctx.saved_tensors
        L = _cabi.lib()
        dev = pts.device
        p, vt, nj = pts.shape[0], lbs_weights.shape[0], lbs_weights.shape[1]
        f32 = torch.float32
        g = g.contiguous().float()
        with torch.cuda.device(dev):
            g_pts = torch.empty((p, 3), dtype=f32, device=dev)
            g_b = torch.zeros((vt, 12), dtype=f32, device=dev)
            g_t = torch.zeros(3, dtype=f32, device=dev) if ctx.has_trans else None
            _cabi.check(L.d3h_lbs_apply_backward(g.data_ptr(), p, idx.data_ptr(), b_inv.data_ptr(), b_pose.data_ptr(), can.data_ptr(),
                                                 pts.data_ptr(), g_pts.data_ptr(), g_b.data_ptr(),
                                                 g_t.data_ptr() if g_t is not None else None, _st(dev)), "d3h_lbs_apply_backward")
            g_a = None
            if ctx.needs_input_grad[1]:
                g_a = torch.empty((nj, 4, 4), dtype=f32, device=dev)
                _cabi.check(L.d3h_lbs_blend_backward(lbs_weights.data_ptr(), g_b.data_ptr(), vt, nj, g_a.data_ptr(), _st(dev)),
                            "d3h_lbs_blend_backward")
        LinearBlendSkinning.launches += 2
        return (g_pts if ctx.needs_input_grad[0] else None, g_a, g_t if ctx.needs_input_grad[2] else None, None, None, None)


class LinearBlendSkinning:
    """vs_template (1,Vt,3) or (Vt,3); lbs_weights (Vt,J); init_A (1,J,4,4) or (J,4,4): what `SMPLX_Deformer.initialize`
    leaves in `self.vs_template`, `self.lbs_weights`, `self.init_A` (:51, :232-235)."""
    launches = 0      # library kernels enqueued so far

    def __init__(self, vs_template, lbs_weights, init_A, k: int = 1):
        if k != 1:
            raise NotImplementedError("d3human-code_b200 skinning: k = 1 nearest template vertex (the reference hard-codes self.k = 1, "
                                      "smplx_exavatar_deformer.py:39)")
        if not vs_template.is_cuda:
            raise RuntimeError("d3human-code_b200 has no CPU path: the skinning tables must live on a CUDA device")
        self.k = 1
        self.vs_template = vs_template.detach().reshape(-1, 3).float().contiguous()
        self.lbs_weights = lbs_weights.detach().float().contiguous()
        self.init_A = init_A.detach().reshape(-1, 4, 4).float().contiguous()
        vt, nj = self.lbs_weights.shape
        if self.vs_template.shape[0] != vt or self.init_A.shape[0] != nj:
            raise ValueError("vs_template / lbs_weights / init_A do not belong to the same rig")
        dev = self.vs_template.device
        self._b_inv = torch.empty((vt, 16), dtype=torch.float32, device=dev)       # (sum_j w[v,j] init_A_j)^-1 per template vertex
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().d3h_lbs_blend(self.lbs_weights.data_ptr(), self.init_A.data_ptr(), vt, nj, 1, self._b_inv.data_ptr(),
                                                  _st(dev)), "d3h_lbs_blend")

    # ---- interpolate_weights (:363-383) -------------------------------------------------------------------------------
    def nearest(self, pts):
        """(P,3) -> (P,) int32: index of the nearest template vertex (knn_points with K = 1)."""
        pts = pts.detach().reshape(-1, 3).float().contiguous()
        L = _cabi.lib()
        dev = pts.device
        p = pts.shape[0]
        idx = torch.empty(p, dtype=torch.int32, device=dev)
        nbytes = int(L.d3h_lbs_nearest_workspace_bytes(p))
        ws = torch.empty(max(nbytes // 4, 16), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(L.d3h_lbs_nearest(pts.data_ptr(), p, self.vs_template.data_ptr(), self.vs_template.shape[0], idx.data_ptr(),
                                          ws.data_ptr(), nbytes, _st(dev)), "d3h_lbs_nearest")
        LinearBlendSkinning.launches += 3
        return idx

    def interpolate_weights(self, pts):
        """(B,P,3) -> (B,P,J) like the reference (with K = 1 the inverse-distance weight is exactly 1)."""
        b, p = pts.shape[0], pts.shape[1]
        return self.lbs_weights[self.nearest(pts).long()].reshape(b, p, -1)

    # ---- lbs_forward_inverse (:424-430) -------------------------------------------------------------------------------
    def lbs_forward_inverse(self, pts):
        """(B,P,3) -> (B,P,3): back to the canonical pose through init_A (no gradient is taken through it in the reference's
        call sites; here it is a plain forward op)."""
        shape = pts.shape
        x = pts.detach().reshape(-1, 3).float().contiguous()
        idx = self.nearest(x)
        can = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().d3h_lbs_apply(x.data_ptr(), x.shape[0], idx.data_ptr(), self._b_inv.data_ptr(), None, None,
                                                  can.data_ptr(), None, _st(x.device)), "d3h_lbs_apply")
        LinearBlendSkinning.launches += 1
        return can.reshape(shape)

    # ---- lbs_forward :472-476 -----------------------------------------------------------------------------------------
    def lbs_transform(self, pts, A, trans=None):
        """pts (1,P,3) or (P,3), A (1,J,4,4) or (J,4,4) = the SMPL-X layer's joint transforms of the frame, trans (1,3) ->
        (P,3) = apply_lbs_inverse(apply_lbs_inverse(pts, init_A, w), A, w, Inverse=False) + trans, reshaped (-1,3) (:476)."""
        x = pts.reshape(-1, 3).float().contiguous()
        a = A.reshape(-1, 4, 4).float().contiguous()
        t = trans.reshape(3).float().contiguous() if trans is not None else None
        idx = self.nearest(x)
        return _TransformFn.apply(x, a, t, idx, self._b_inv, self.lbs_weights)
