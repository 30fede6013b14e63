"""Comparison helpers shared by the parity tests."""
import numpy as np

# north_star: H estimates and combined symbols within 1e-5 relative error (fp32), bits exact.
REL_TOL = 1e-5


def rel_errors(got, ref):
    """(max-norm, 2-norm) error relative to the reference's max / 2-norm, per SURVEY Appendix A.11."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    d = got.astype(np.complex128) - ref.astype(np.complex128)
    mx = float(np.abs(d).max() / max(np.abs(ref).max(), 1e-30))
    l2 = float(np.linalg.norm(d.ravel()) / max(np.linalg.norm(ref.astype(np.complex128).ravel()), 1e-30))
    return mx, l2


def assert_close(got, ref, what, tol=REL_TOL):
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(np.asarray(got).view(np.float32)).all(), f"{what}: non-finite values"
    mx, l2 = rel_errors(got, ref)
    assert mx <= tol and l2 <= tol, f"{what}: max-rel {mx:.3e}, l2-rel {l2:.3e} > {tol:g}"
    return mx, l2


def threshold_margin(sym, qam_bits):
    """smallest distance of any combined symbol component to a demap decision threshold"""
    s = np.asarray(sym)
    comps = np.concatenate([s.real.ravel(), s.imag.ravel()]).astype(np.float64)
    d = np.abs(comps)
    if qam_bits == 4:
        d = np.minimum(d, np.abs(np.abs(comps) - 2 / np.sqrt(10)))
    elif qam_bits == 6:
        for t in (2 / np.sqrt(42), 4 / np.sqrt(42), 6 / np.sqrt(42)):
            d = np.minimum(d, np.abs(np.abs(comps) - t))
    return float(d.min())
