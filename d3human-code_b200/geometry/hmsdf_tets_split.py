"""`hmSDF_Tets`: drop-in for the reference class of the same name (geometry/hmsdf_tets_split.py:89-454).

Same algorithm as `GShell_Tets` plus the `type` argument (hmsdf_tets_split.py:254,261-264):
    type == "cloth": msdf_n is used as is;  type == "body": -msdf_n;  anything else: unchanged.
Faithful to the reference, type == "body" does not back-propagate into `msdf_n`: the negation happens under
torch.no_grad() there (:256-264), so msdf_n.grad stays None while pos / sdf still receive gradients.
"""
from __future__ import annotations

from ..extract import extract, extract_frames
from .gshell_tets import _TetsExtractor


class hmSDF_Tets(_TetsExtractor):
    def __call__(self, pos_nx3, sdf_n, msdf_n, tet_fx4, type, output_watertight_template=True):
        return extract(pos_nx3, sdf_n, msdf_n, tet_fx4, msdf_negate=(type == "body"),
                       output_watertight_template=output_watertight_template)

    def split(self, pos_nx3, sdf_n, msdf_n, tet_fx4, output_watertight_template=True, fused=False):
        """Both extractions of one split-stage iteration in one call (extension; the reference calls the class twice,
        train.py:1040-1047 via hmsdf.py:548): returns `(cloth_tuple, body_tuple)`, each the reference's 6-tuple, exactly
        what `self(..., "cloth")` and `self(..., "body")` return; gradients of `pos_nx3` / `sdf_n` are the sum over both,
        `msdf_n` only receives the cloth part (the reference negates msdf under no_grad for the body,
        hmsdf_tets_split.py:256-264).
        fused=True (EXPERIMENTAL, see extract_frames_async): the pair shares one classification / edge de-duplication /
        vertex interpolation and only the mSDF cut runs twice (with output_watertight_template=False the two do not
        share their valid tets, gshell_tets.py:275, and the plain batch is used)."""
        if not (fused and output_watertight_template):
            # measured on a B200 (bench.py `split_pair`, configs[2] grid): two lean single calls (0.41 ms per pair fwd+bwd)
            # beat the generic two-frame batch (0.47 ms) -- the pair is too small to amortise the batch machinery
            return (extract(pos_nx3, sdf_n, msdf_n, tet_fx4, False, output_watertight_template),
                    extract(pos_nx3, sdf_n, msdf_n, tet_fx4, True, output_watertight_template))
        cloth, body = extract_frames([pos_nx3, pos_nx3], sdf_n, msdf_n, tet_fx4, types=["cloth", "body"],
                                     output_watertight_template=output_watertight_template, lanes=2, fused_pair=True)
        return cloth, body
