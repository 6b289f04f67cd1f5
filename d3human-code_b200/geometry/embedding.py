"""`Embedding`: drop-in for the reference module of the same name (geometry/embedding.py:4-38): x -> (x, sin(2^k x),
cos(2^k x), ...), k < N_freqs.  The arithmetic runs in libd3h_tets.so (d3h_mlp_embed); CUDA tensors only."""
from __future__ import annotations

import torch
from torch import nn

from .. import _cabi


class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, n_freq):
        m = x.shape[0]
        c = 3 * (2 * n_freq + 1)
        out = torch.empty((m, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().d3h_mlp_embed(x.data_ptr(), m, n_freq, out.data_ptr(), c, c,
                                                  torch.cuda.current_stream(x.device).cuda_stream), "d3h_mlp_embed")
        ctx.save_for_backward(x)
        ctx.n_freq = n_freq
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous().float()
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().d3h_mlp_embed_backward(x.data_ptr(), x.shape[0], ctx.n_freq, g.data_ptr(), g.shape[1],
                                                           gx.data_ptr(), 0, torch.cuda.current_stream(x.device).cuda_stream),
                        "d3h_mlp_embed_backward")
        return gx, None


class Embedding(nn.Module):
    def __init__(self, in_channels, N_freqs, logscale=True):
        super().__init__()
        self.N_freqs = N_freqs
        self.in_channels = in_channels
        self.funcs = [torch.sin, torch.cos]
        self.out_channels = in_channels * (len(self.funcs) * N_freqs + 1)
        if logscale:
            self.freq_bands = 2 ** torch.linspace(0, N_freqs - 1, N_freqs)
        else:
            self.freq_bands = torch.linspace(1, 2 ** (N_freqs - 1), N_freqs)
        self.logscale = logscale

    def forward(self, x):
        if self.in_channels != 3 or not self.logscale:
            raise NotImplementedError("d3human-code_b200 Embedding: 3 input channels and logscale=True (what D3-Human uses)")
        if not x.is_cuda:
            raise RuntimeError("d3human-code_b200 has no CPU path: Embedding needs a CUDA tensor")
        shape = x.shape
        y = _EmbedFn.apply(x.reshape(-1, 3).contiguous().float(), self.N_freqs)
        return y.reshape(*shape[:-1], self.out_channels)
