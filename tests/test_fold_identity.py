"""CPU: the index algebra the fold / FFT-pass kernels rely on, checked in numpy with hypothesis.

For an every-s-th-bin far-field grid, the M-point DFT of the fftshifted aperture restricted to bins that
are multiples of s equals a K = M/s point DFT of the s-fold aliased aperture, with the fftshift of the
input turned into a roll of the folded aperture and the fftshift of the output into a roll of the K-point
spectrum (csrc/fold.cu, csrc/fft.cu; reference nearfield_farfield.py:18-20, :68)."""
import numpy as np
from hypothesis import given, settings, strategies as st


def fold_then_dft(J, s1, s2):
    """What mlb_fft_rows (fold + rolls) followed by mlb_fft_cols computes, in numpy."""
    M1, M2 = J.shape
    K1, K2 = M1 // s1, M2 // s2
    h1, h2 = M1 // 2, M2 // 2
    p1 = (np.arange(K1) - h1) % K1
    p2 = (np.arange(K2) - h2) % K2
    G = np.zeros((K1, K2), complex)
    for t1 in range(s1):
        for t2 in range(s2):
            G += J[np.ix_(p1 + t1 * K1, p2 + t2 * K2)]          # G[p] = sum_t J[((p-h) mod K) + tK]
    F = np.fft.fft2(G)
    return np.roll(F, ((h1 // s1) % K1, (h2 // s2) % K2), axis=(0, 1))   # out[(q + h/s) mod K] = F[q]


@settings(max_examples=40, deadline=None)
@given(st.sampled_from([(8, 1), (8, 2), (8, 4), (12, 2), (12, 3), (16, 4), (20, 2), (24, 4), (36, 6), (10, 5)]),
       st.sampled_from([(8, 1), (8, 4), (12, 2), (16, 8), (18, 3), (20, 10), (6, 1)]), st.integers(0, 2 ** 31 - 1))
def test_fold_identity(ms1, ms2, seed):
    (M1, s1), (M2, s2) = ms1, ms2
    assert M1 % s1 == 0 and (M1 // 2) % s1 == 0 and M2 % s2 == 0 and (M2 // 2) % s2 == 0
    rng = np.random.default_rng(seed)
    J = rng.standard_normal((M1, M2)) + 1j * rng.standard_normal((M1, M2))
    ref = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(J)))[::s1, ::s2]
    got = fold_then_dft(J, s1, s2)
    assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()


@settings(max_examples=25, deadline=None)
@given(st.integers(3, 40), st.integers(3, 40), st.integers(0, 2 ** 31 - 1))
def test_dense_twiddles_reproduce_fft_for_any_size(M1, M2, seed):
    """Separable dense form with phase origin M - M//2 equals fft2(fftshift(.)) for odd,
    even and non-square apertures -- the contract of FarfieldPlan(method='dense')."""
    rng = np.random.default_rng(seed)
    J = rng.standard_normal((M1, M2)) + 1j * rng.standard_normal((M1, M2))
    def tw(M):
        o = M - M // 2
        q = (np.arange(M) - M // 2) % M                    # un-shifted bin number of shifted index
        return np.exp(-2j * np.pi * np.outer(np.arange(M) - o, q) / M)      # [m][q']
    F = tw(M1).T @ J @ tw(M2)
    ref = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(J)))
    assert np.abs(F - ref).max() <= 1e-10 * np.abs(ref).max()


def test_tile_fold_conditions_match_plan_rules():
    """FarfieldPlan allows fold/fft only when the stride divides M and M//2 (otherwise the strided bins are
    not multiples of s and a residual modulation would be needed)."""
    for M, s, ok in ((4096, 4, True), (128, 4, True), (12, 4, False), (10, 5, True), (18, 4, False), (48, 4, True)):
        assert ((M % s == 0) and ((M // 2) % s == 0)) == ok
