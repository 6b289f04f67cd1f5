#!/usr/bin/env python
"""SDF field query on one GPU: the D3-Human network (n_freq 6, 256 wide, 6 hidden layers, skip_in [3]) on
all N grid vertices of the 128^3 grid in batches of 100000 points (hmsdf.py:187, 434-444), forward and forward+backward,
this package (tcgen05, 3xTF32) against the same module written with plain PyTorch ops in fp32 (how the reference runs),
and with torch's TF32 switch on (lower precision than the reference: shown for scale only).  Prints one JSON line.
    python profiles/mlp_bench.py [--points 2146689] [--batch 100000] [--reps 5]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from torch import nn

CFG = dict(n_freq=6, d_hidden=256, n_hidden=6, skip_in=[3])      # train.py:1622-1625


class TorchMLP(nn.Module):     # the op sequence of geometry/mlp.py:33-45 + embedding.py:23-38 on the same parameters
    def __init__(self, src):
        super().__init__()
        self.net, self.skip_count, self.n_freq = src.net, src.skip_count, src._plan.n_freq

    def forward(self, x):
        out = [x]
        for k in range(self.n_freq):
            out += [torch.sin(2.0 ** k * x), torch.cos(2.0 ** k * x)]
        emb = torch.cat(out, -1)
        x = emb
        for i, m in enumerate(self.net):
            x = m(torch.cat([x, emb], -1)) if i in self.skip_count else m(x)
        return x


def measure(points=129 ** 3, batch=100000, reps=5, dev=None, tf32_leg=True):
    from d3human_code_b200.geometry import mlp as M
    dev = dev or torch.device("cuda:0")
    torch.manual_seed(0)
    ours = M.MLP(**CFG).to(dev)
    ref = TorchMLP(ours)
    pts = (torch.rand(points, 3, device=dev) * 2 - 1)
    flops_fwd = 2.0 * (39 * 256 + 5 * 256 * 256 + 295 * 256 + 256) * points

    def run(net, backward):
        for p in ours.parameters():
            p.grad = None
        sdf = torch.cat([net(pts[i:i + batch]) for i in range(0, points, batch)], 0)
        if backward:
            sdf.sum().backward()
        return sdf

    def timed(net, backward):
        run(net, backward)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(net, backward); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    res = {"points": points, "batch": batch, "network": CFG, "flops_fwd": flops_fwd,
           "unit": "ms for all points (median of %d, CUDA events)" % reps}
    with torch.no_grad():
        y0, y1 = run(ours, False), run(ref, False)
    res["max_abs_diff_vs_torch_fp32"] = float((y0 - y1).abs().max())
    res["max_abs_sdf"] = float(y1.abs().max())
    l0 = M.launch_counter()
    for name, net in (("ours", ours), ("torch_fp32", ref)):
        with torch.no_grad():
            res[name + "_fwd_ms"] = timed(net, False)
        res[name + "_fwd_bwd_ms"] = timed(net, True)
    res["gpu_launches"] = M.launch_counter() - l0
    if tf32_leg:
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            with torch.no_grad():
                res["torch_tf32_fwd_ms"] = timed(ref, False)
                res["torch_tf32_max_abs_diff_vs_fp32"] = float((run(ref, False) - y1).abs().max())
            res["torch_tf32_fwd_bwd_ms"] = timed(ref, True)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
    res["ours_fwd_tflops_fp32_equiv"] = flops_fwd / (res["ours_fwd_ms"] * 1e-3) / 1e12
    res["ours_fwd_tensor_tflops"] = 3 * res["ours_fwd_tflops_fp32_equiv"]
    res["ours_fwd_bwd_tensor_tflops"] = 9 * flops_fwd / (res["ours_fwd_bwd_ms"] * 1e-3) / 1e12
    res["points_per_s_fwd_bwd"] = points / (res["ours_fwd_bwd_ms"] * 1e-3)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=129 ** 3)
    ap.add_argument("--batch", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(measure(a.points, a.batch, a.reps)))
