#!/usr/bin/env python
"""Device time of every building block of the SDF-network stage at the reference's batch size (M = 100000 points):
median CUDA-event time of 20 calls each.  python profiles/mlp_blocks_timing.py [M]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from d3human_code_b200 import _cabi

m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
dev = torch.device("cuda:0")
L = _cabi.lib()
st = torch.cuda.current_stream(dev).cuda_stream
x = torch.rand(m, 3, device=dev) * 2 - 1
a256 = torch.randn(m, 256, device=dev)
a320 = torch.randn(m, 320, device=dev)
dz = torch.randn(m, 256, device=dev)
y = torch.rand(m, 256, device=dev) * 0.05
out = torch.empty(m, 320, device=dev)
w = torch.randn(256, 320, device=dev) / 16
bias = torch.randn(256, device=dev) * 0.1
wp = torch.empty(2 * 256 * 320, device=dev)
dw = torch.zeros(256, 320, device=dev)
db = torch.zeros(256, device=dev)
nb = int(L.d3h_mlp_wgrad_workspace_bytes(m, 256, 256))
ws = torch.empty(nb // 4, device=dev)
g = torch.randn(m, 1, device=dev)
w1 = torch.randn(1, 256, device=dev)
dw1, db1 = torch.zeros(1, 256, device=dev), torch.zeros(1, device=dev)
sdf = torch.empty(m, 1, device=dev)


def t(name, fn, flops=None, bytes_=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:7.1f} TFLOP/s (fp32-equivalent; x3 on the tensor core)"
    if bytes_:
        extra += f"  {bytes_ / us / 1e3:7.1f} GB/s"
    print(f"{name:34s} {us:8.1f} us{extra}")


pk = lambda n, k: _cabi.check(L.d3h_mlp_pack_weight(w.data_ptr(), 320, n, k, 0, 0, 0, n, k, wp.data_ptr(), st), "pack")
pk(256, 256)
t("pack_weight 256x256", lambda: pk(256, 256))
t("embed (39 -> 64 columns)", lambda: L.d3h_mlp_embed(x.data_ptr(), m, 6, out.data_ptr(), 64, 64, st), bytes_=m * (12 + 256))
t("linear K=256 N=256 softplus", lambda: L.d3h_mlp_linear(a256.data_ptr(), 256, m, 256, wp.data_ptr(), 256, bias.data_ptr(), 1, None, 0, out.data_ptr(), 256, st), flops=2.0 * m * 256 * 256, bytes_=m * 2048)
t("linear K=256 N=256 identity", lambda: L.d3h_mlp_linear(a256.data_ptr(), 256, m, 256, wp.data_ptr(), 256, bias.data_ptr(), 0, None, 0, out.data_ptr(), 256, st), flops=2.0 * m * 256 * 256)
t("linear K=256 N=256 mode 2 (dX)", lambda: L.d3h_mlp_linear(dz.data_ptr(), 256, m, 256, wp.data_ptr(), 256, None, 2, y.data_ptr(), 256, out.data_ptr(), 256, st), flops=2.0 * m * 256 * 256, bytes_=m * 3072)
pk(256, 320)
t("linear K=320 N=256 softplus", lambda: L.d3h_mlp_linear(a320.data_ptr(), 320, m, 320, wp.data_ptr(), 256, bias.data_ptr(), 1, None, 0, out.data_ptr(), 256, st), flops=2.0 * m * 320 * 256)
pk(256, 64)
t("linear K=64 N=256 softplus", lambda: L.d3h_mlp_linear(a256.data_ptr(), 256, m, 64, wp.data_ptr(), 256, bias.data_ptr(), 1, None, 0, out.data_ptr(), 256, st), flops=2.0 * m * 64 * 256)
pk(64, 256)
t("linear K=256 N=64 identity", lambda: L.d3h_mlp_linear(dz.data_ptr(), 256, m, 256, wp.data_ptr(), 64, None, 0, None, 0, out.data_ptr(), 64, st), flops=2.0 * m * 64 * 256)
t("wgrad N=256 K=256", lambda: L.d3h_mlp_wgrad(dz.data_ptr(), 256, a256.data_ptr(), 256, m, 256, 256, dw.data_ptr(), 320, db.data_ptr(), ws.data_ptr(), nb, st), flops=2.0 * m * 256 * 256, bytes_=m * 2048)
t("wgrad N=256 K=64", lambda: L.d3h_mlp_wgrad(dz.data_ptr(), 256, a256.data_ptr(), 256, m, 256, 64, dw.data_ptr(), 320, None, ws.data_ptr(), nb, st), flops=2.0 * m * 64 * 256)
t("head (256 -> 1)", lambda: L.d3h_mlp_head(a256.data_ptr(), 256, m, 256, w1.data_ptr(), db1.data_ptr(), 1, sdf.data_ptr(), st), bytes_=m * 1024)
t("head_backward", lambda: L.d3h_mlp_head_backward(a256.data_ptr(), 256, m, 256, w1.data_ptr(), 1, g.data_ptr(), out.data_ptr(), 256, dw1.data_ptr(), db1.data_ptr(), st), bytes_=m * 2048)
t("embed_backward", lambda: L.d3h_mlp_embed_backward(x.data_ptr(), m, 6, a256.data_ptr(), 256, out.data_ptr(), 0, st), bytes_=m * (12 + 256 + 12))
t("torch fp32 matmul 256x256 (cuBLAS)", lambda: torch.matmul(a256, w[:, :256].t()), flops=2.0 * m * 256 * 256)
t("torch zeros (256,320)", lambda: torch.zeros(256, 320, device=dev))
