"""Parity metric shared by the tests: P spans many decades and
contains NaN, so compare  max|dP| / max|P_ref|  over finite bins AND require the
NaN (evanescent) masks to be identical."""
import numpy as np

FF_TOL = 1e-5      # north_star: fp32 outputs within 1e-5 relative of the reference


def power_map_error(P, P_ref):
    P = np.asarray(P, dtype=np.float64)
    P_ref = np.asarray(P_ref, dtype=np.float64)
    assert P.shape == P_ref.shape, (P.shape, P_ref.shape)
    nan_a, nan_b = np.isnan(P), np.isnan(P_ref)
    assert np.array_equal(nan_a, nan_b), "NaN (evanescent) masks differ: %d vs %d" % (nan_a.sum(), nan_b.sum())
    fin = ~nan_b
    scale = np.abs(P_ref[fin]).max()
    if scale == 0:
        return float(np.abs(P[fin]).max())
    return float(np.abs(P[fin] - P_ref[fin]).max() / scale)


def field_error(F, F_ref):
    F = np.asarray(F).astype(np.complex128)
    F_ref = np.asarray(F_ref).astype(np.complex128)
    scale = np.abs(F_ref).max()
    return float(np.abs(F - F_ref).max() / (scale if scale else 1.0))
