"""shared helpers for the parity tests (test infrastructure)."""
import numpy as np

# north_star: floating-point outputs and gradients within 1e-5 relative in fp32
TOL = 1e-5
# stated looser bound of the opt-in single-pass TF32 tap contraction
TOL_TF32 = 5e-3
# stated bound of the opt-in single fp16 plane mode of the tcgen05 wide path (include/gfc.h: GFC_PREC_F16)
TOL_F16 = 2e-3


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = max(float(np.abs(ref).max()), 1e-30)
    return float(np.abs(got - ref).max()) / den


def assert_close(got, ref, tol=TOL, name=""):
    """SURVEY §8c rule: max|d| <= tol*max|ref| per tensor AND allclose(rtol=tol, atol=tol*max|ref|)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (name, got.shape, ref.shape)
    assert np.isfinite(got).all(), "%s: non-finite values" % name
    e = rel_err(got, ref)
    assert e <= tol, "%s: normwise error %.3e > %.1e" % (name, e, tol)
    scale = max(float(np.abs(ref).max()), 1e-30)
    assert np.allclose(got, ref, rtol=tol, atol=tol * scale), "%s: allclose(rtol=atol=%g) failed" % (name, tol)
    return e
