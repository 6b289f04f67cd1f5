"""Build libd3h_tets.so (hand-written sm_100a kernels + the C ABI of include/d3h_tets.h) in-tree with nvcc.

    python d3human-code_b200/build.py [--force]

-fmad=false: the float pipeline must round like the reference's separate PyTorch kernels; the few
places that need a fused multiply-add (torch.cross, torch.linspace restatements) call __fmaf_rn explicitly.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libd3h_tets.so")
SOURCES = ["d3h_api.cu", "d3h_classify.cu", "d3h_sort.cu", "d3h_scan.cu", "d3h_surface.cu", "d3h_backward.cu", "d3h_mesh.cu", "d3h_mlp.cu", "d3h_lbs.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--threads", "0"]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", h) for h in ("d3h_tets.h", "d3h_mesh.h", "d3h_mlp.h", "d3h_lbs.h")] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libd3h_tets.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(os.path.join(os.path.dirname(LIB), "ptxas_info.txt"), "w") as fh:
        fh.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
