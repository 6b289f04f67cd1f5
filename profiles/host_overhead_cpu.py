#!/usr/bin/env python
"""Pure host cost of the Python path, measured WITHOUT a GPU: the library is replaced by a stub that only publishes
fixed sizes (no compute), tensors live on the CPU, so what is timed is extract.py + autograd bookkeeping.
    python profiles/host_overhead_cpu.py [--frames 8] [--iters 300] [--profile]"""
import argparse, contextlib, cProfile, ctypes as C, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import _cabi
from d3human_code_b200 import extract as E

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--iters", type=int, default=300)
ap.add_argument("--n-grid", type=int, default=35937)       # 33^3: small tensors, the Python cost does not depend on N
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
SIZES = dict(fv=5774, t1=3894, t2=1880, v=3827, fa=4191)      # a tenth of the 128^3 surface


class NullLib:
    def d3h_version(self): return _cabi.VERSION
    def d3h_last_error_string(self): return b""
    def d3h_workspace_bytes_static(self, f, n, cap, ne): return 4096
    def d3h_lanes_join(self, s): return 0
    def d3h_wait_counts(self, p, seq, t): return 0
    def d3h_extract_backward_batch(self, p, n, l, s): return 0
    def d3h_extract_forward(self, ptr, stream): return self.d3h_extract_forward_batch_nojoin(ptr, 1, 1, stream)
    def d3h_extract_backward(self, ptr, stream): return 0
    def d3h_extract_forward_batch_nojoin(self, ptr, n, lanes, stream):
        size = C.sizeof(_cabi.ForwardArgs)
        for i in range(n):
            a = _cabi.ForwardArgs.from_address(int(ptr) + i * size)
            c = _cabi.Counts.from_address(int(a.counts_host))
            c.n_valid_tets, c.n_tri_tets, c.n_quad_tets = SIZES["fv"], SIZES["t1"], SIZES["t2"]
            c.n_corners = 3 * SIZES["t1"] + 4 * SIZES["t2"]
            c.n_verts, c.n_faces_aug, c.seq = SIZES["v"], SIZES["fa"], a.seq
        return 0


class _Stream:
    cuda_stream = 0


_cabi.lib = lambda: NullLib()
E._check_cuda = lambda t: None
E.packed_tets = lambda t, n: t
torch.cuda.current_stream = lambda dev=None: _Stream()
torch.cuda.device = lambda dev=None: contextlib.nullcontext()
torch.Tensor.pin_memory = lambda self: self
E.set_static_edges("0")

N, B = args.n_grid, args.frames
tets = torch.zeros((6 * 32 ** 3, 4), dtype=torch.int32)
pos1 = torch.zeros((N, 3), requires_grad=True)
posb = torch.zeros((B, N, 3), requires_grad=True)
sdf = torch.zeros((N, 1), requires_grad=True)
msdf = torch.zeros(N, requires_grad=True)


def single():
    pos1.grad = sdf.grad = msdf.grad = None
    verts, faces, _, _, _, extra = E.extract(pos1, sdf, msdf, tets)
    torch.autograd.backward([verts, extra["msdf"]], [gv1, gm1])


def batch():
    posb.grad = sdf.grad = msdf.grad = None
    outs = E.extract_frames_async(posb, sdf, msdf, tets, types="cloth", lanes=8).result()
    torch.autograd.backward([o[0] for o in outs] + [o[5]["msdf"] for o in outs], gvb + gmb)


o = E.extract(pos1, sdf, msdf, tets)
gv1, gm1 = torch.zeros_like(o[0]), torch.zeros_like(o[5]["msdf"])
ob = E.extract_frames(posb, sdf, msdf, tets, types="cloth")
gvb, gmb = [torch.zeros_like(x[0]) for x in ob], [torch.zeros_like(x[5]["msdf"]) for x in ob]
def batch_packed():
    posb.grad = sdf.grad = msdf.grad = None
    pk = E.extract_frames_async(posb, sdf, msdf, tets, types="cloth", lanes=8).packed()
    torch.autograd.backward([pk.verts_aug, pk.msdf], [gvp, gmp])


def single_generic():
    pos1.grad = sdf.grad = msdf.grad = None
    verts, faces, _, _, _, extra = E.extract_generic(pos1, sdf, msdf, tets)
    torch.autograd.backward([verts, extra["msdf"]], [gv1, gm1])


pk0 = E.extract_frames_async(posb, sdf, msdf, tets, types="cloth", lanes=8).packed()
gvp, gmp = torch.zeros_like(pk0.verts_aug), torch.zeros_like(pk0.msdf)
for fn, name, per in ((batch_packed, f"packed batch of {B} frames fwd+bwd", B), (single, "drop-in single call fwd+bwd", 1), (single_generic, "  (through the batch machinery)", 1), (batch, f"batch of {B} frames fwd+bwd", B)):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        fn()
    dt = (time.perf_counter() - t0) / args.iters
    print(f"{name:34s} {dt * 1e6:8.1f} us per call   {dt * 1e6 / per:7.1f} us per frame")
if args.profile:
    for fn in (single, batch):
        pr = cProfile.Profile(); pr.enable()
        for _ in range(args.iters):
            fn()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)
