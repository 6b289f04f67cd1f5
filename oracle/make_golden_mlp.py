"""Writes tests/golden/mlp_*.npz from the LIVE reference module (`/root/reference/geometry/mlp.py`, fp32, CPU): inputs,
parameters, outputs and the autograd gradients of a scalar loss.  Run in the build container only:

    python oracle/make_golden_mlp.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference_mlp  # noqa: E402

CASES = {   # name: (n_freq, d_hidden, n_hidden, skip_in, n_points, seed)
    "mlp_d3human": (6, 256, 6, [3], 300, 0),       # train.py:1622-1626: the configuration D3-Human trains
    "mlp_default": (6, 128, 3, [], 257, 1),        # the class defaults (mlp.py:10)
    "mlp_two_skips": (4, 64, 4, [0, 2], 130, 2),
}


def main():
    ref = load_reference_mlp()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (n_freq, d_hidden, n_hidden, skip_in, m, seed) in CASES.items():
        torch.manual_seed(seed)
        net = ref.MLP(n_freq=n_freq, d_hidden=d_hidden, d_out=1, n_hidden=n_hidden, skip_in=skip_in)
        with torch.no_grad():      # softplus(beta=100) saturates on default-init weights: spread the pre-activations
            for p in net.parameters():
                p.mul_(1.5)
        x = (torch.rand(m, 3) * 2 - 1).requires_grad_(True)
        y = net(x)
        gy = torch.randn_like(y)
        (y * gy).sum().backward()
        lin = [mod for mod in net.net if isinstance(mod, torch.nn.Linear)]
        data = dict(x=x.detach().numpy(), y=y.detach().numpy(), gy=gy.numpy(), gx=x.grad.numpy(),
                    cfg=np.array([n_freq, d_hidden, n_hidden, m], dtype=np.int64), skip_in=np.array(skip_in, dtype=np.int64))
        for i, l in enumerate(lin):
            data[f"w{i}"], data[f"b{i}"] = l.weight.detach().numpy(), l.bias.detach().numpy()
            data[f"gw{i}"], data[f"gb{i}"] = l.weight.grad.numpy(), l.bias.grad.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **data)
        print(name, "y range", float(y.min()), float(y.max()))


if __name__ == "__main__":
    main()
