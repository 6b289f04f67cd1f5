"""Plain-PyTorch restatement of the extraction path, for TIMING the way the reference runs today on a GPU.

TEST / BENCH INFRASTRUCTURE ONLY (same rules as the rest of oracle/).  The reference implements this path as ~200
PyTorch ops on device='cuda' (geometry/gshell_tets.py:253-447); /root/reference does not exist on the GPU box, so the
"reference-style PyTorch on the same B200" figure that SURVEY.md section 8(d) asks for is measured with this port: the
same kinds of operations in the same places -- boolean-mask compactions (one host sync each), `torch.unique(dim=0,
return_inverse=True)` over the 6*Fv sorted edge rows (:279), gathers through the case tables, scatter_add splats for
normals and tangents (:9-78) including the 4*ceil(sqrt(F))^2-row UV table that map_uv materialises (:219-233), masked
boundary interpolation (:342-385), the six-bucket cut (:400-420) and the unused-row zeroing (:423-427) -- with autograd
providing the backward pass.  Written from SURVEY.md Appendix A and oracle/gshell_oracle.py, not from the reference's text.

Pinned by tests/test_oracle_torch_port.py: integer outputs equal the golden vectors of the live reference exactly,
positions / mSDF to 1e-6, gradients to 1e-5.  It is NOT the parity oracle (that is gshell_oracle.py); `kind` is "port".
"""
from __future__ import annotations

import math

import torch

from . import gshell_oracle as O


def _tables(dev):
    t = lambda a: torch.as_tensor(a, device=dev)  # noqa: E731
    return dict(tri=t(O.TRIANGLE_TABLE), loop=t(O.MESH_EDGE_TABLE), cut3=t(O.TRIANGLE_TABLE_TRI), cut4=t(O.TRIANGLE_TABLE_QUAD),
                n=t(O.NUM_TRIANGLES_TABLE), n3=t(O.NUM_TRIANGLES_TRI_TABLE), n4=t(O.NUM_TRIANGLES_QUAD_TABLE),
                edges=t(O.BASE_TET_EDGES))


def _dot(a, b):
    return (a * b).sum(-1, keepdim=True)


def _safe_normalize(x):
    return x / torch.sqrt(torch.clamp(_dot(x, x), min=1e-20))


def _uv_table(num_tets, dev):
    """The per-tet UV atlas (4 corners per cell of an n x n grid), built in full like map_uv does."""
    n = int(math.ceil(math.sqrt((num_tets * 2 + 1) // 2)))
    lin = torch.linspace(0, 1 - (1 / n), n, dtype=torch.float32, device=dev)
    ty, tx = torch.meshgrid(lin, lin, indexing="ij")
    pad = 0.9 / n
    corners = [torch.stack([tx + dx, ty + dy], -1) for dx, dy in ((0.0, 0.0), (pad, 0.0), (pad, pad), (0.0, pad))]
    return torch.stack(corners, -2).reshape(-1, 2)


def _auto_normals(verts, faces):
    i0, i1, i2 = faces[:, 0], faces[:, 1], faces[:, 2]
    v0, v1, v2 = verts[i0], verts[i1], verts[i2]
    if faces.shape[0] == 3:
        fn = torch.cross(v1 - v0, v2 - v0, dim=0)  # what torch.cross without `dim` does on (3,3) operands
    else:
        fn = torch.cross(v1 - v0, v2 - v0, dim=-1)
    nrm = torch.zeros_like(verts)
    for i in (i0, i1, i2):
        nrm = nrm.scatter_add(0, i[:, None].repeat(1, 3), fn)
    nrm = torch.where(_dot(nrm, nrm) > 1e-20, nrm, torch.tensor([0.0, 0.0, 1.0], device=verts.device))
    return _safe_normalize(nrm)


def _tangents(verts, v_nrm, faces, uvs):
    pos = [verts[faces[:, i]] for i in range(3)]
    tex = [uvs[faces[:, i]] for i in range(3)]  # quirk: the position indices double as texture indices
    uve1, uve2 = tex[1] - tex[0], tex[2] - tex[0]
    pe1, pe2 = pos[1] - pos[0], pos[2] - pos[0]
    nom = pe1 * uve2[..., 1:2] - pe2 * uve1[..., 1:2]
    den = uve1[..., 0:1] * uve2[..., 1:2] - uve1[..., 1:2] * uve2[..., 0:1]
    tang = nom / torch.where(den > 0.0, torch.clamp(den, min=1e-6), torch.clamp(den, max=-1e-6))
    tangents = torch.zeros_like(v_nrm)
    count = torch.zeros_like(v_nrm)
    for i in range(3):
        idx = faces[:, i][:, None].repeat(1, 3)
        tangents = tangents.scatter_add(0, idx, tang)
        count = count.scatter_add(0, idx, torch.ones_like(tang))
    tangents = _safe_normalize(tangents / count)
    return _safe_normalize(tangents - _dot(tangents, v_nrm) * v_nrm)


def extract(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_sign: int = 1, output_watertight_template: bool = True):
    """-> (verts_aug, faces_aug, None, None, v_tng_aug, extra): the reference's tuple, built with plain torch ops."""
    dev = pos_nx3.device
    T = _tables(dev)
    s = sdf_n.float().reshape(-1)
    if msdf_sign < 0:
        with torch.no_grad():  # hmsdf_tets_split.py:261-264 negates without a graph
            m = -msdf_n
    else:
        m = msdf_n
    tets = tet_fx4.long()

    with torch.no_grad():
        occ = s > 0
        occ_fx4 = occ[tets.reshape(-1)].reshape(-1, 4)
        occ_sum = occ_fx4.sum(-1)
        valid = (occ_sum > 0) & (occ_sum < 4)
        if not output_watertight_template:
            valid = valid & ((m > 0)[tets.reshape(-1)].reshape(-1, 4).sum(-1) > 0)
        tv = tets[valid]
        all_edges = tv[:, T["edges"]].reshape(-1, 2)
        all_edges = torch.sort(all_edges, dim=1).values
        unique_edges, idx_map = torch.unique(all_edges, dim=0, return_inverse=True)
        mask_edges = occ[unique_edges.reshape(-1)].reshape(-1, 2).sum(-1) == 1
        mapping = torch.full((unique_edges.shape[0],), -1, dtype=torch.long, device=dev)
        n_verts = int(mask_edges.sum())
        mapping[mask_edges] = torch.arange(n_verts, dtype=torch.long, device=dev)
        idx_map = mapping[idx_map].reshape(-1, 6)
        interp_v = unique_edges[mask_edges]
        code = (occ_fx4[valid].long() * torch.tensor([1, 2, 4, 8], device=dev)).sum(-1)
        ntri = T["n"][code]
    ea, eb = interp_v[:, 0], interp_v[:, 1]

    # zero-crossing interpolation
    e0, e1 = s[ea], -s[eb]
    d = e0 + e1
    dd = torch.sign(d) * (torch.abs(d) + 1e-12)
    dd = torch.where(dd == 0, torch.full_like(dd, 1e-12), dd)
    w0, w1 = e1 / dd, e0 / dd
    verts = pos_nx3[ea] * w0[:, None] + pos_nx3[eb] * w1[:, None]
    msdf_vert = m[ea] * w0 + m[eb] * w1
    msdf_vert_sg = m[ea] * w0.detach() + m[eb] * w1.detach()  # the returned attribute: weights without a graph

    # watertight faces
    is1, is2 = ntri == 1, ntri == 2
    faces = torch.cat([
        torch.gather(idx_map[is1], 1, T["tri"][code[is1]][:, :3]).reshape(-1, 3),
        torch.gather(idx_map[is2], 1, T["tri"][code[is2]][:, :6]).reshape(-1, 3)], 0)
    uvs = _uv_table(tets.shape[0], dev)
    v_nrm = _auto_normals(verts, faces)
    v_tng = _tangents(verts, v_nrm, faces, uvs)

    # polygon loops and boundary vertices (one per polygon edge)
    loop3 = torch.gather(idx_map[is1], 1, T["loop"][code[is1]][:, :3])
    loop4 = torch.gather(idx_map[is2], 1, T["loop"][code[is2]][:, :4])
    cur = torch.cat([loop3.reshape(-1), loop4.reshape(-1)])
    nxt = torch.cat([torch.roll(loop3, -1, 1).reshape(-1), torch.roll(loop4, -1, 1).reshape(-1)])
    mi, mj = msdf_vert[cur], msdf_vert[nxt]
    big_d = mi + (-mj)
    nz = (torch.abs(torch.sign(mi) + torch.sign(mj)) != 2) & (torch.abs(big_d) > 1e-12)
    safe_d = torch.where(nz, big_d, torch.ones_like(big_d))
    u0 = torch.where(nz, (-mj) / safe_d, torch.zeros_like(big_d))
    u1 = torch.where(nz, mi / safe_d, torch.zeros_like(big_d))
    bverts = verts[cur] * u0[:, None] + verts[nxt] * u1[:, None]
    btng = v_tng[cur] * u0[:, None] + v_tng[nxt] * u1[:, None]
    bmsdf = msdf_vert_sg[cur] * u0.detach() + msdf_vert_sg[nxt] * u1.detach()
    verts_aug = torch.cat([verts, bverts], 0)
    v_tng_aug = torch.cat([v_tng, btng], 0)
    msdf_aug = torch.cat([msdf_vert_sg, bmsdf], 0)

    # mSDF cut in six buckets
    with torch.no_grad():
        t1, t2 = loop3.shape[0], loop4.shape[0]
        case3 = ((msdf_vert[loop3] > 0).long() * torch.tensor([4, 2, 1], device=dev)).sum(-1)
        case4 = ((msdf_vert[loop4] > 0).long() * torch.tensor([8, 4, 2, 1], device=dev)).sum(-1)
        loc3 = torch.cat([loop3, n_verts + torch.arange(3 * t1, device=dev).reshape(-1, 3)], 1)
        loc4 = torch.cat([loop4, n_verts + 3 * t1 + torch.arange(4 * t2, device=dev).reshape(-1, 4)], 1)
        n3, n4 = T["n3"][case3], T["n4"][case4]
        buckets = []
        for k in (1, 2):
            sel = n3 == k
            buckets.append(torch.gather(loc3[sel], 1, T["cut3"][case3[sel]][:, :3 * k]).reshape(-1, 3))
        for k in (1, 2, 3, 4):
            sel = n4 == k
            buckets.append(torch.gather(loc4[sel], 1, T["cut4"][case4[sel]][:, :3 * k]).reshape(-1, 3))
        faces_aug = torch.cat(buckets, 0)
        used = torch.zeros(verts_aug.shape[0], dtype=torch.bool, device=dev)
        used[faces_aug.unique()] = True
    verts_aug = verts_aug * used[:, None]  # unreferenced rows are zeroed (and receive no gradient)

    extra = {"msdf": msdf_aug, "msdf_watertight": msdf_vert_sg, "msdf_boundary": msdf_aug[n_verts:]}
    if output_watertight_template:
        extra = dict(n_verts_watertight=n_verts, vertices_watertight=verts, faces_watertight=faces,
                     v_tng_watertight=v_tng, **extra)
    return verts_aug, faces_aug, None, None, v_tng_aug, extra
