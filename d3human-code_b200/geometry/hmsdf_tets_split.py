"""`hmSDF_Tets`: drop-in for the reference class of the same name (geometry/hmsdf_tets_split.py:89-454).

Same algorithm as `GShell_Tets` plus the `type` argument (hmsdf_tets_split.py:254,261-264):
    type == "cloth": msdf_n is used as is;  type == "body": -msdf_n;  anything else: unchanged.
Faithful to the reference, type == "body" does not back-propagate into `msdf_n`: the negation happens under
torch.no_grad() there (:256-264), so msdf_n.grad stays None while pos / sdf still receive gradients.
"""
from __future__ import annotations

from ..extract import extract
from .gshell_tets import _TetsExtractor


class hmSDF_Tets(_TetsExtractor):
    def __call__(self, pos_nx3, sdf_n, msdf_n, tet_fx4, type, output_watertight_template=True):
        return extract(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_negate=(type == "body"),
                       output_watertight_template=output_watertight_template)
