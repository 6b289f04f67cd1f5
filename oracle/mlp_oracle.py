"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs).

CPU restatement, in numpy float64, of the SDF field query that feeds the extraction:

    Embedding.forward      geometry/embedding.py:23-38   out = [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]
    MLP.__init__/forward   geometry/mlp.py:10-45         Linear + Softplus(beta=100) chain; before the Linear of hidden
                                                         layer i in skip_in the input becomes cat([x, emb], -1) (:41)
    nn.Softplus(beta=100)  threshold 20 (torch default): y = z if 100 z > 20 else log1p(exp(100 z)) / 100

`forward` keeps the layer inputs; `backward` is the hand-derived adjoint (gradients of every weight / bias and of the
query points).  Pinned against the live reference module (fp32) in tests/test_mlp_oracle.py and against the golden
vectors tests/golden/mlp_*.npz written from the live reference by oracle/make_golden_mlp.py.
"""
from __future__ import annotations

import numpy as np

BETA, THRESHOLD = 100.0, 20.0


def embed(x, n_freq):
    """geometry/embedding.py:23-38 (logscale=True: freq_bands = 2 ** linspace(0, n_freq - 1, n_freq))."""
    x = np.asarray(x, dtype=np.float64)
    out = [x]
    for k in range(n_freq):
        f = 2.0 ** k
        out += [np.sin(f * x), np.cos(f * x)]
    return np.concatenate(out, -1)


def embed_backward(x, n_freq, g_emb):
    x = np.asarray(x, dtype=np.float64)
    g = g_emb[:, 0:3].copy()
    for k in range(n_freq):
        f = 2.0 ** k
        g += f * (np.cos(f * x) * g_emb[:, 3 + 6 * k:6 + 6 * k] - np.sin(f * x) * g_emb[:, 6 + 6 * k:9 + 6 * k])
    return g


def softplus(z):
    t = BETA * z
    return np.where(t > THRESHOLD, z, np.log1p(np.exp(np.minimum(t, THRESHOLD))) / BETA)


def softplus_grad(z):
    t = BETA * z
    return np.where(t > THRESHOLD, 1.0, 1.0 / (1.0 + np.exp(-np.minimum(t, THRESHOLD))))


def layer_shapes(n_freq=6, d_hidden=128, d_out=1, n_hidden=3, skip_in=()):
    """(fan_out, fan_in) of the Linear layers in MLP.net order (mlp.py:13-31)."""
    e = 3 * (2 * n_freq + 1)
    shapes = [(d_hidden, e)]
    for i in range(n_hidden):
        shapes.append((d_hidden, d_hidden + e) if i in skip_in else (d_hidden, d_hidden))
    shapes.append((d_out, d_hidden))
    return shapes


def forward(x, weights, biases, n_freq, skip_in=()):
    """x (M,3); weights / biases in MLP.net order (first, n_hidden hidden ones, output).  -> (out (M, d_out), cache)."""
    emb = embed(x, n_freq)
    w = [np.asarray(a, dtype=np.float64) for a in weights]
    b = [np.asarray(a, dtype=np.float64) for a in biases]
    inputs, pre = [], []
    h = emb
    n_hidden = len(w) - 2
    for li in range(len(w)):
        if 1 <= li <= n_hidden and (li - 1) in skip_in:
            h = np.concatenate([h, emb], -1)         # mlp.py:41: cat([x, emb])
        inputs.append(h)
        z = h @ w[li].T + b[li]
        pre.append(z)
        h = softplus(z) if li < len(w) - 1 else z
    cache = dict(x=np.asarray(x, dtype=np.float64), emb=emb, inputs=inputs, pre=pre, w=w, n_freq=n_freq, skip_in=tuple(skip_in))
    return h, cache


def backward(cache, g_out):
    """-> (g_x (M,3), [g_weight...], [g_bias...]) for upstream g_out (M, d_out)."""
    w, inputs, pre = cache["w"], cache["inputs"], cache["pre"]
    n_layers = len(w)
    n_hidden = n_layers - 2
    e = cache["emb"].shape[1]
    gw, gb = [None] * n_layers, [None] * n_layers
    g_emb = np.zeros_like(cache["emb"])
    g = np.asarray(g_out, dtype=np.float64)          # gradient at the pre-activation of the output layer
    for li in range(n_layers - 1, -1, -1):
        gw[li] = g.T @ inputs[li]
        gb[li] = g.sum(0)
        g_in = g @ w[li]
        if 1 <= li <= n_hidden and (li - 1) in cache["skip_in"]:
            g_emb += g_in[:, -e:]
            g_in = g_in[:, :-e]
        if li == 0:
            g_emb += g_in
        else:
            g = g_in * softplus_grad(pre[li - 1])
    return embed_backward(cache["x"], cache["n_freq"], g_emb), gw, gb
