"""Writes tests/golden/lbs_*.npz from the reference's own skinning methods (deform/smplx_exavatar_deformer.py:363-421,
executed from the reference source; see oracle/ref_loader.load_reference_lbs_methods).  Build container only:
    python oracle/make_golden_lbs.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_lbs_oracle import reference_lbs, synthetic_rig  # noqa: E402


def main():
    for name, seed, vt, j, p in (("lbs_small", 0, 300, 8, 500), ("lbs_smplx_like", 1, 2000, 55, 3000)):
        template, w, init_a, a, trans = synthetic_rig(seed, vt, j)
        rng = np.random.default_rng(seed)
        pts = rng.uniform(-1.1, 1.1, size=(p, 3)).astype(np.float32)
        pts[::4] = 0.0
        g = rng.standard_normal(pts.shape).astype(np.float32)
        posed, can, g_pts, g_a, g_t = reference_lbs(pts, template, w, init_a, a, trans, g)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), pts=pts, template=template, w=w, init_a=init_a,
                            a=a, trans=trans, g=g, posed=posed, canonical=can, g_pts=g_pts, g_a=g_a, g_trans=g_t)
        print(name, posed.shape)


if __name__ == "__main__":
    main()
