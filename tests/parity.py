"""Parity checker used by the GPU tests: the north_star tolerance is 1e-9 relative per time step for state and discharge."""
import numpy as np

RTOL = 1.0e-9


def parity_report(got, want, rtol=RTOL, atol_frac=1.0e-13):
    """-> (n_bad, worst_rel, index_of_worst).  |got-want| <= rtol*max(|got|,|want|) + atol_frac*max|want| ; NaN must match NaN."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    scale = np.nanmax(np.abs(want)) if np.any(~nan_w) else 0.0
    diff = np.abs(got - want)
    tol = rtol * np.maximum(np.abs(got), np.abs(want)) + atol_frac * scale
    bad = (diff > tol) & ~(nan_g & nan_w)
    bad |= nan_g ^ nan_w
    rel = np.where(nan_g | nan_w, 0.0, diff / np.maximum(np.maximum(np.abs(got), np.abs(want)), 1e-300))
    rel = np.where(diff <= atol_frac * scale, 0.0, rel)
    worst = np.unravel_index(np.argmax(rel), rel.shape) if rel.size else ()
    return int(bad.sum()), float(rel.max()) if rel.size else 0.0, worst


def assert_parity(got, want, name, rtol=RTOL, atol_frac=1.0e-13, max_bad=0):
    n_bad, worst, where = parity_report(got, want, rtol, atol_frac)
    assert n_bad <= max_bad, f"{name}: {n_bad} of {np.size(want)} values outside rtol {rtol:g}; worst rel {worst:.3e} at {where}: " \
                             f"got {np.asarray(got)[where]!r} want {np.asarray(want)[where]!r}"
    return worst
