"""Shared comparison helpers for the parity tests (oracle vs golden, CUDA vs oracle, CUDA vs golden)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
#: tolerances stated by BASELINE.json north_star
POS_RTOL = 1e-6      # interpolated positions (we are bit-exact in practice; the assert uses exact first)
GRAD_RTOL = 1e-5     # gradients, normwise (max |diff| / max |ref|)
TNG_ATOL = 2e-5      # unit tangents, oracle vs reference on CPU: same scatter order, ATen's approximate sqrt -> few ulp
#: CUDA vs oracle: the per-face tangents are ~1/denominator large (UV cells are 1/ceil(sqrt(F)) apart) and are summed
#: with float atomics in arbitrary order (the reference's own CUDA scatter_add_ is order-nondeterministic too), so the
#: normalised result differs by eps * sum|t_f| / |sum t_f|: bound the worst row loosely and the bulk tightly.
TNG_CUDA_ATOL = 5e-3
TNG_CUDA_P99 = 1e-4


def golden_cases():
    """extraction fixtures (oracle/make_golden.py); the mesh_* files belong to oracle/make_golden_mesh.py, the mlp_* files
    to oracle/make_golden_mlp.py"""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(("mesh_", "mlp_", "lbs_"))]


def golden_path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    rec = {k: z[k] for k in z.files}
    cls, typ, wt = [str(x) for x in rec.pop("meta")]
    rec["cls"], rec["type"], rec["wt"] = cls, (None if typ == "None" else typ), wt == "True"
    rec["sign"] = -1 if rec["type"] == "body" else 1
    return rec


def assert_exact(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{name}: dtype {got.dtype} != {want.dtype}"
    if not np.array_equal(got, want):
        bad = np.nonzero(got != want)
        raise AssertionError(f"{name}: {len(bad[0])} mismatching elements, first at {[b[0] for b in bad]}: "
                             f"{got[tuple(b[0] for b in bad)]} != {want[tuple(b[0] for b in bad)]}")


def assert_close_normwise(name, got, want, rtol):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    if want.size == 0:
        return
    scale = np.abs(want).max()
    err = np.abs(got - want).max()
    assert err <= rtol * max(scale, 1e-30), f"{name}: max|diff|={err:.3e} > {rtol:g} * max|ref|={scale:.3e}"


def assert_tangents_close(name, got, want, atol=TNG_ATOL, p99=None):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype
    both_nan = np.isnan(got) & np.isnan(want)
    diff = np.where(both_nan, 0.0, np.abs(got.astype(np.float64) - want.astype(np.float64)))
    assert not np.isnan(diff).any(), f"{name}: NaN pattern differs"
    assert diff.max(initial=0.0) <= atol, f"{name}: max|diff|={diff.max():.3e} > {atol:g}"
    if p99 is not None and diff.size:
        q = np.quantile(diff.max(-1), 0.99)
        assert q <= p99, f"{name}: 99th percentile row error {q:.3e} > {p99:g}"


#: condition-aware tangent check (CUDA vs oracle): a row may differ by TNG_BASE + TNG_PER_KAPPA * kappa, kappa from
#: oracle.tangent_condition (1.3e-7 = one fp32 ulp of a unit vector component, re-ordered sums of <= ~12 terms); rows whose
#: bound reaches TNG_UNCHECKED are ill-conditioned in the reference itself and only have to be finite
TNG_BASE = 2e-6
TNG_PER_KAPPA = 4e-7
TNG_UNCHECKED = 0.05


def assert_tangents_conditioned(name, got, want, kappa, max_unchecked_frac=None):
    """|got - want| <= TNG_BASE + TNG_PER_KAPPA * kappa per row; returns the fraction of rows too ill-conditioned to
    check.  `max_unchecked_frac` (smooth fields) bounds that fraction so that the check cannot become vacuous."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype
    if got.shape[0] == 0:
        return 0.0
    bound = TNG_BASE + TNG_PER_KAPPA * np.asarray(kappa, np.float64)
    checked = bound < TNG_UNCHECKED
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64)).max(-1)
    bad = checked & ~(diff <= bound)          # NaN in a checked row fails
    if bad.any():
        i = int(np.argmax(np.where(bad, diff / bound, 0)))
        raise AssertionError(f"{name}: {int(bad.sum())} rows exceed the condition-aware bound; worst row {i}: "
                             f"|diff|={diff[i]:.3e} > {bound[i]:.3e} (kappa={kappa[i]:.3e})")
    finite = np.isfinite(got).all(-1) | ~np.isfinite(want).all(-1)
    assert finite.all(), f"{name}: non-finite rows where the oracle is finite"
    frac = 1.0 - checked.mean()
    if max_unchecked_frac is not None:
        assert frac <= max_unchecked_frac, f"{name}: {frac:.3f} of the rows are too ill-conditioned to check"
    return frac


def check_forward_against_golden(out, rec, tangents=True, tng_atol=TNG_ATOL, tng_p99=None, kappa=None):
    """out: dict with the reference's return values (numpy). rec: golden record.  kappa: per-row tangent condition
    (oracle.tangent_condition) -> the condition-aware tangent check replaces the flat tolerance."""
    assert_exact("faces_aug", out["faces_aug"], rec["faces_aug"])
    assert_exact("verts_aug", out["verts_aug"], rec["verts_aug"])
    assert_exact("msdf", out["msdf"], rec["extra_msdf"])
    assert_exact("msdf_watertight", out["msdf_watertight"], rec["extra_msdf_watertight"])
    assert_exact("msdf_boundary", out["msdf_boundary"], rec["extra_msdf_boundary"])
    if rec["wt"]:
        assert int(out["n_verts_watertight"]) == int(rec["extra_n_verts_watertight"])
        assert_exact("faces_watertight", out["faces_watertight"], rec["extra_faces_watertight"])
        assert_exact("vertices_watertight", out["vertices_watertight"], rec["extra_vertices_watertight"])
        if tangents and kappa is not None:
            nv = int(rec["extra_n_verts_watertight"])
            assert_tangents_conditioned("v_tng_watertight", out["v_tng_watertight"], rec["extra_v_tng_watertight"], kappa[:nv])
        elif tangents:
            assert_tangents_close("v_tng_watertight", out["v_tng_watertight"], rec["extra_v_tng_watertight"], tng_atol, tng_p99)
    else:
        assert "extra_vertices_watertight" not in rec
    if tangents and kappa is not None:
        assert_tangents_conditioned("v_tng_aug", out["v_tng_aug"], rec["v_tng_aug"], kappa)
    elif tangents:
        assert_tangents_close("v_tng_aug", out["v_tng_aug"], rec["v_tng_aug"], tng_atol, tng_p99)


def check_grads_against_golden(g_pos, g_sdf, g_msdf, rec):
    assert_close_normwise("grad_pos", g_pos, rec["grad_pos"], GRAD_RTOL)
    assert_close_normwise("grad_sdf", np.asarray(g_sdf).reshape(rec["grad_sdf"].shape), rec["grad_sdf"], GRAD_RTOL)
    if "grad_msdf" in rec:
        assert g_msdf is not None
        assert_close_normwise("grad_msdf", g_msdf, rec["grad_msdf"], GRAD_RTOL)
    else:
        assert g_msdf is None, "type='body' must not produce an msdf gradient (hmsdf_tets_split.py:261-264)"
