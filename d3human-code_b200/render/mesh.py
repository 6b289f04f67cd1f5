"""Drop-in for the two pieces of the reference's `render/mesh.py` that run on every extracted surface:

    Mesh            render/mesh.py:139-250   container whose constructor computes the distinct face edges (get_edge :240-250)
    auto_normals    render/mesh.py:418-446   smooth vertex normals, differentiable w.r.t. v_pos

`geometry/hmsdf.py` builds three to five `Mesh` objects per extraction and runs `auto_normals` on each
(hmsdf.py:460-484, 554-593).  Both are served by sm_100a kernels behind the C ABI of include/d3h_mesh.h
(csrc/d3h_mesh.cu); there is no PyTorch / CPU fallback.

Differences by design, results being the reference's:
  * `Mesh.edges` is computed on FIRST ACCESS from the faces the mesh was built with, not in the constructor.  Nothing in
    D3-Human reads the edges of the extracted meshes, so the torch.unique(dim=0) the reference pays in every
    constructor is never run.  Meshes that share one faces tensor share one edge list (cache on data_ptr / version).
  * a face index outside [0, V) raises IndexError when the edges are read (the reference would return the edge).
  * normals are computed in fp32 whatever the dtype of v_pos and cast back.
  * two `auto_normals` calls on the same (v_pos, t_pos_idx) tensor objects return the SAME normals tensor while the
    first result is alive (hmsdf.py:558-561 and :589-593 do exactly that every iteration); D3H_MESH_SHARE_NORMALS=0
    turns it off for callers that modify normals in place.
The rest of the reference module (OBJ loading, AABB helpers, tangent space, Laplacian) is outside this row.
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, Tuple

import torch
from torch.autograd.function import once_differentiable

from .. import _cabi

_FIELDS = ("v_pos", "t_pos_idx", "v_nrm", "t_nrm_idx", "v_tex", "t_tex_idx", "v_tng", "t_tng_idx", "material", "kd", "ks",
           "uv", "uv_idx", "face_labels", "v_labels", "connected_faces")
_TENSOR_FIELDS = tuple(f for f in _FIELDS if f != "material")  # what clone() copies (render/mesh.py:203-236)

_COUNT_SLOTS = 64
_CW = 4  # int64 words of d3h_mesh_counts
_launches = 0  # kernels launched by this module (edges: 5 + 2 fills, normals: 2 + 1 fill, adjoint: 1 + 1 fill)


def launch_counter() -> int:
    return _launches


def _check_cuda(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError("d3human-code_b200 has no CPU path: mesh tensors must live on a CUDA device "
                           "(the reference hard-codes device='cuda' as well, render/mesh.py:439)")


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _DeviceState:
    """Scratch of one (device, stream): the edge workspace and a ring of pinned, device-mapped size slots."""

    def __init__(self, device):
        self.device = device
        self.workspace = None
        self.workspace_bytes = 0
        self.counts_host = torch.zeros(_COUNT_SLOTS * _CW, dtype=torch.int64).pin_memory()
        self.counts_np = self.counts_host.numpy().reshape(_COUNT_SLOTS, _CW)
        self.counts_dev = torch.zeros(_CW, dtype=torch.int64, device=device)
        self.seq = 0

    def ensure(self, need: int) -> int:
        if need > self.workspace_bytes:
            raw = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 256
            self.workspace = raw[off:off + need]
            self.workspace_bytes = need
        return self.workspace.data_ptr()


_states: Dict[Tuple, _DeviceState] = {}


def _state_for(device) -> _DeviceState:
    key = (device.type, device.index, _stream(device))
    st = _states.get(key)
    if st is None:
        st = _states[key] = _DeviceState(device)
    return st


def reset() -> None:
    """Drop cached workspaces and edge lists (tests)."""
    _states.clear()
    _edge_cache.clear()
    del _normals_memo[:]


# ---- edges --------------------------------------------------------------------------------------------------------
_edge_cache: Dict[Tuple, Tuple[torch.Tensor, "weakref.ref"]] = {}


def face_edges(t_pos_idx: torch.Tensor, n_verts: int) -> torch.Tensor:
    """Distinct undirected edges of (F,3) faces over `n_verts` vertices -> (E,2) int64, rows in ascending lexicographic
    order: what `Mesh.get_edge` returns (render/mesh.py:240-250).  One size read from mapped host memory, no stream
    synchronisation."""
    global _launches
    _check_cuda(t_pos_idx)
    if t_pos_idx.dim() != 2 or t_pos_idx.shape[1] != 3:
        raise ValueError(f"t_pos_idx must have shape (F,3), got {tuple(t_pos_idx.shape)}")
    faces = t_pos_idx.to(torch.int64).contiguous()
    n_faces = faces.shape[0]
    if n_faces == 0:
        return faces.new_zeros((0, 2))
    L = _cabi.lib()
    dev = faces.device
    with torch.cuda.device(dev):
        st = _state_for(dev)
        need = L.d3h_mesh_edges_workspace_bytes(n_faces, n_verts)
        if need < 0:
            raise RuntimeError(f"mesh too large for d3h_mesh_edges: F={n_faces}, V={n_verts}")
        ws_ptr = st.ensure(need)
        cap = 3 * n_faces
        edges = torch.empty((cap, 2), dtype=torch.int64, device=dev)
        st.seq += 1
        slot = st.seq % _COUNT_SLOTS
        host_ptr = st.counts_host.data_ptr() + slot * _CW * 8
        _cabi.check(L.d3h_mesh_edges(faces.data_ptr(), n_faces, n_verts, edges.data_ptr(), cap, ws_ptr, st.workspace_bytes,
                                     st.counts_dev.data_ptr(), host_ptr, st.seq, _stream(dev)), "d3h_mesh_edges")
        _launches += 7
        _cabi.check(L.d3h_mesh_wait_counts(host_ptr, st.seq, 60_000_000), "d3h_mesh_wait_counts")
        n_edges, bad, overflow, _ = (int(x) for x in st.counts_np[slot])
    if bad:
        raise IndexError(f"t_pos_idx holds vertex indices outside [0, {n_verts})")
    if overflow:
        raise RuntimeError("d3h_mesh_edges: edge capacity overflow (cannot happen with cap = 3F)")
    return edges[:n_edges]


def _edges_cached(faces: torch.Tensor, n_verts: int) -> torch.Tensor:
    key = (faces.data_ptr(), faces._version, tuple(faces.shape), faces.dtype, faces.device, int(n_verts))
    hit = _edge_cache.get(key)
    if hit is not None and hit[1]() is not None:
        return hit[0]
    out = face_edges(faces, n_verts)
    if len(_edge_cache) > 64:
        for k in [k for k, v in _edge_cache.items() if v[1]() is None]:
            del _edge_cache[k]
    _edge_cache[key] = (out, weakref.ref(faces))
    return out


# ---- the container -------------------------------------------------------------------------------------------------
class Mesh:
    """Same constructor signature and attributes as the reference (render/mesh.py:139-163)."""

    def __init__(self, v_pos=None, t_pos_idx=None, v_nrm=None, t_nrm_idx=None, v_tex=None, t_tex_idx=None, v_tng=None,
                 t_tng_idx=None, material=None, base=None, kd=None, ks=None, uv=None, uv_idx=None, face_labels=None,
                 v_labels=None, connected_faces=None, edges=None):
        given = dict(v_pos=v_pos, t_pos_idx=t_pos_idx, v_nrm=v_nrm, t_nrm_idx=t_nrm_idx, v_tex=v_tex, t_tex_idx=t_tex_idx,
                     v_tng=v_tng, t_tng_idx=t_tng_idx, material=material, kd=kd, ks=ks, uv=uv, uv_idx=uv_idx,
                     face_labels=face_labels, v_labels=v_labels, connected_faces=connected_faces)
        for name in _FIELDS:
            setattr(self, name, given[name])
        if base is not None:
            self.copy_none(base)
        if self.t_pos_idx is None:
            # the reference's constructor always runs get_edge (:162), which fails on a mesh without faces
            raise TypeError("Mesh needs t_pos_idx (directly or through `base`): 'NoneType' object is not subscriptable")
        # the reference overwrites whatever `edges` was passed or inherited with get_edge() (:162); here the same list is
        # produced when somebody asks for it, from the faces the mesh holds NOW (later re-assignments do not change it)
        self._edge_faces = self.t_pos_idx
        self._edge_verts = None if self.v_pos is None else int(self.v_pos.shape[0])
        self._edges = None

    def copy_none(self, other):
        """render/mesh.py:167-201 (the edge list is always recomputed by the constructor, so it is not copied)."""
        for name in _FIELDS:
            if getattr(self, name) is None:
                setattr(self, name, getattr(other, name))

    def clone(self):
        """render/mesh.py:203-238: a detached deep copy of every tensor attribute."""
        out = Mesh(base=self)
        for name in _TENSOR_FIELDS:
            val = getattr(out, name)
            if val is not None:
                setattr(out, name, val.clone().detach())
        out._edge_faces = out.t_pos_idx
        if self._edges is not None:
            out._edges = self._edges.clone().detach()
        return out

    @property
    def edges(self):
        if self._edges is None:
            self.get_edge()
        return self._edges

    @edges.setter
    def edges(self, value):
        self._edges = value

    def get_edge(self):
        """render/mesh.py:240-250."""
        faces = self._edge_faces
        n_verts = self._edge_verts
        if n_verts is None:  # a mesh without positions: the index range is all there is (one host read)
            n_verts = int(faces.max().item()) + 1 if faces.numel() else 0
        self._edges = _edges_cached(faces, n_verts)
        return self._edges


def accelerate(reference_mesh_cls):
    """For a maintainer who keeps the reference's own `render/mesh.py` (OBJ loading, AABB helpers, Laplacian, ...): returns
    `(FastMesh, auto_normals)` where FastMesh is a SUBCLASS of the reference's Mesh (render/mesh.py:139) -- every method and
    attribute of the original stays -- whose constructor no longer runs torch.unique (the edge list is produced by the
    kernels on first access of `.edges`), and an `auto_normals` that builds FastMesh objects.  At the bottom of the
    reference module:

        from d3human_code_b200.render.mesh import accelerate
        Mesh, auto_normals = accelerate(Mesh)
    """

    class FastMesh(reference_mesh_cls):
        def __init__(self, *args, **kwargs):
            self._d3h_constructing = True
            self._edges = None
            super().__init__(*args, **kwargs)     # ends with self.get_edge(), which only takes note while constructing
            self._d3h_constructing = False

        @property
        def edges(self):
            if self._edges is None and not self._d3h_constructing and getattr(self, "_edge_faces", None) is not None:
                self.get_edge()
            return self._edges

        @edges.setter
        def edges(self, value):
            self._edges = value

        def get_edge(self):
            if self._d3h_constructing:            # called by the reference constructor (:162): remember what to use
                self._edge_faces = self.t_pos_idx
                self._edge_verts = None if self.v_pos is None else int(self.v_pos.shape[0])
                self._edges = None
                return None
            faces, n_verts = self._edge_faces, self._edge_verts
            if n_verts is None:
                n_verts = int(faces.max().item()) + 1 if faces.numel() else 0
            self._edges = _edges_cached(faces, n_verts)
            return self._edges

    FastMesh.__name__ = getattr(reference_mesh_cls, "__name__", "Mesh")
    FastMesh.__qualname__ = FastMesh.__name__

    def fast_auto_normals(imesh):
        v_nrm = vertex_normals(imesh.v_pos, imesh.t_pos_idx)
        if torch.is_anomaly_enabled():
            assert torch.all(torch.isfinite(v_nrm))
        return FastMesh(v_nrm=v_nrm, t_nrm_idx=imesh.t_pos_idx, base=imesh)

    return FastMesh, fast_auto_normals


# ---- normals -------------------------------------------------------------------------------------------------------
class _AutoNormalsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_pos, t_pos_idx):
        global _launches
        L = _cabi.lib()
        dev = v_pos.device
        pos = v_pos.detach().float().contiguous()
        faces = t_pos_idx.to(torch.int64).contiguous()
        n_verts, n_faces = pos.shape[0], faces.shape[0]
        v_nrm = torch.empty((n_verts, 3), dtype=torch.float32, device=dev)
        acc = torch.empty((n_verts, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(L.d3h_mesh_normals_forward(pos.data_ptr(), faces.data_ptr(), n_verts, n_faces, v_nrm.data_ptr(),
                                                   acc.data_ptr(), None, _stream(dev)), "d3h_mesh_normals_forward")
        _launches += 3
        ctx.save_for_backward(pos, faces, acc)
        ctx.in_dtype = v_pos.dtype
        return v_nrm.to(v_pos.dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_nrm):
        global _launches
        pos, faces, acc = ctx.saved_tensors
        L = _cabi.lib()
        dev = pos.device
        g = g_nrm.float().contiguous()
        g_pos = torch.empty_like(pos)
        with torch.cuda.device(dev):
            _cabi.check(L.d3h_mesh_normals_backward(pos.data_ptr(), faces.data_ptr(), pos.shape[0], faces.shape[0],
                                                    acc.data_ptr(), g.data_ptr(), g_pos.data_ptr(), _stream(dev)),
                        "d3h_mesh_normals_backward")
        _launches += 2
        return g_pos.to(ctx.in_dtype), None


# `getMesh_split` runs auto_normals twice on the very same (verts, faces) pair (hmsdf.py:558-561 and :589-593).  The
# result of the last few calls is remembered through weak references: a second call with the same two tensor objects
# (unchanged versions, same grad mode) returns the same normals tensor while somebody still holds it.
_normals_memo: list = []
_NORMALS_MEMO = 4
share_normals = os.environ.get("D3H_MESH_SHARE_NORMALS", "1") != "0"


def _memo_lookup(v_pos, faces, needs_grad):
    for ent in _normals_memo:
        rp, vp, rf, vf, ro, vo = ent
        out = ro()
        if (out is not None and rp() is v_pos and rf() is faces and vp == v_pos._version and vf == faces._version
                and vo == out._version and out.requires_grad == needs_grad):
            return out
    return None


def vertex_normals(v_pos: torch.Tensor, t_pos_idx: torch.Tensor) -> torch.Tensor:
    """(V,3) positions, (F,3) faces -> (V,3) unit normals (render/mesh.py:420-441); differentiable w.r.t. v_pos."""
    _check_cuda(v_pos)
    _check_cuda(t_pos_idx)
    if v_pos.dim() != 2 or v_pos.shape[1] != 3:
        raise ValueError(f"v_pos must have shape (V,3), got {tuple(v_pos.shape)}")
    if t_pos_idx.dim() != 2 or t_pos_idx.shape[1] != 3:
        raise ValueError(f"t_pos_idx must have shape (F,3), got {tuple(t_pos_idx.shape)}")
    if not share_normals:
        return _AutoNormalsFn.apply(v_pos, t_pos_idx)
    needs_grad = torch.is_grad_enabled() and v_pos.requires_grad
    out = _memo_lookup(v_pos, t_pos_idx, needs_grad)
    if out is None:
        out = _AutoNormalsFn.apply(v_pos, t_pos_idx)
        _normals_memo.insert(0, (weakref.ref(v_pos), v_pos._version, weakref.ref(t_pos_idx), t_pos_idx._version,
                                 weakref.ref(out), out._version))
        del _normals_memo[_NORMALS_MEMO:]
    return out


def auto_normals(imesh: Mesh) -> Mesh:
    """render/mesh.py:418-446."""
    v_nrm = vertex_normals(imesh.v_pos, imesh.t_pos_idx)
    if torch.is_anomaly_enabled():
        assert torch.all(torch.isfinite(v_nrm))
    return Mesh(v_nrm=v_nrm, t_nrm_idx=imesh.t_pos_idx, base=imesh)
