"""CPU: the chirp-z decomposition behind method='czt' (csrc/czt.cu), emulated in numpy with the kernels' exact table
layout, equals the direct sum  F[i] = sum_m J[m] e^{-ik x_m ux_i}  (nearfield_farfield.py:97-120) on uniform grids that are
NOT FFT bins."""
import numpy as np
import pytest


def chirps(M, o, K, L, t_lin, t_quad):
    """pre / kern / post as czt_chirps_kernel builds them."""
    mp = np.arange(M) - o
    pre = np.exp(-1j * np.pi * (t_lin * mp + 0.5 * t_quad * mp * mp))
    kern = np.zeros(L, complex)
    for t in range(L):
        if t <= K - 1 + o:
            n = t
        elif t - L >= o - M + 1:
            n = t - L
        else:
            continue
        kern[t] = np.exp(1j * np.pi * 0.5 * t_quad * n * n)
    post = np.exp(-1j * np.pi * 0.5 * t_quad * np.arange(K) ** 2) / L
    return pre, kern, post


def czt_1d(J, o, K, L, t_lin, t_quad):
    M = J.shape[-1]
    pre, kern, post = chirps(M, o, K, L, t_lin, t_quad)
    a = np.zeros(J.shape[:-1] + (L,), complex)
    a[..., :M] = J * pre
    y = np.conj(np.fft.fft(a, axis=-1) * np.fft.fft(kern))          # the pointwise kernel conjugates ...
    conv = np.conj(np.fft.fft(y, axis=-1))                           # ... so that a FORWARD pass inverts (1/L sits in post)
    return conv[..., o:o + K] * post


@pytest.mark.parametrize("M,K,u0,du", [(48, 20, -0.31, 0.013), (64, 64, 0.05, -0.004), (37, 11, 0.0, 0.02)])
def test_chirp_z_equals_direct_sum(M, K, u0, du):
    rng = np.random.default_rng(M + K)
    wl, n, d = 532e-9, 1.4607, 532e-9 / 2.2
    J = rng.standard_normal((5, M)) + 1j * rng.standard_normal((5, M))
    o = M - M // 2
    u = u0 + du * np.arange(K)
    x = (np.arange(M) - o) * d
    k = 2 * np.pi * n / wl
    direct = J @ np.exp(-1j * k * np.outer(x, u))
    L = 1
    while L < M + K - 1:
        L *= 2
    got = czt_1d(J, o, K, L, 2 * n * d * u0 / wl, 2 * n * d * du / wl)
    assert np.abs(got - direct).max() <= 1e-11 * np.abs(direct).max()
