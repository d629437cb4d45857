"""CPU: the algebra behind the fused column-pass + power kernel (csrc/fft.cu, fft_cols_power_kernel).

The kernel never stores the four aperture sums: per far-field point it keeps only
    t1 = L_phi + Z N_theta,   t2 = L_theta - Z N_phi        (nearfield_farfield.py:158-167, :184)
accumulated field by field as real-coefficient combinations
    t1 = -px Fex - py Fey + Z cy Fhx - Z cx Fhy,   t2 = -cy Fex + cx Fey - Z px Fhx - Z py Fhy
with (px,py) = (ux,uy)/(sin(theta)+1e-9), (cx,cy) = (px,py) uz and the DC bin as (px,py,cx,cy) = (1,0,1,0).
This test checks, in float64 numpy, that P built that way equals the oracle's radiated_power()
(restating :135-189) including the DC special case and the NaN mask of evanescent bins."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import farfield_oracle as fo


def power_from_projections(F, ux, uy, dxp, dyp, wavelength, n_glass):
    ux = np.asarray(ux, float).reshape(-1, 1)
    uy = np.asarray(uy, float).reshape(1, -1)
    uz2 = 1 - ux ** 2 - uy ** 2
    uz = np.sqrt(np.where(uz2 < 0, np.nan, uz2))
    inv = 1.0 / (np.sqrt(ux ** 2 + uy ** 2) + 1e-9)
    dc = (ux == 0) & (uy == 0)
    px = np.where(dc, 1.0, ux * inv)
    py = uy * inv + 0 * ux
    cx, cy = px * uz, py * uz
    Z = fo.Z0 / n_glass
    coef = [(-px, -cy), (-py, cx), (Z * cy, -Z * px), (-Z * cx, -Z * py)]      # Ex, Ey, Hx, Hy
    t1 = sum(k1 * f for (k1, _), f in zip(coef, F))
    t2 = sum(k2 * f for (_, k2), f in zip(coef, F))
    k = 2 * math.pi * n_glass / wavelength
    scale = k * k / (32 * math.pi ** 2 * Z) * (dxp * dyp) ** 2 * 2.0
    return scale * (np.abs(t1) ** 2 + np.abs(t2) ** 2) / (uz + 1e-5)


@settings(max_examples=30, deadline=None)
@given(st.integers(2, 24), st.integers(2, 24), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_projection_form_equals_reference_epilogue(kx, ky, seed, fft_grid):
    rng = np.random.default_rng(seed)
    wl, ng, d = 532e-9, 1.4607, 532e-9 / 2.2
    if fft_grid:        # the reference's own grid: contains the exact DC bin and evanescent corners
        ux = np.fft.fftshift(fo.fft_bin_direction_cosines(kx, d, wl, ng))
        uy = np.fft.fftshift(fo.fft_bin_direction_cosines(ky, d, wl, ng))
    else:
        ux = np.sort(rng.uniform(-1.1, 1.1, kx))
        uy = np.sort(rng.uniform(-1.1, 1.1, ky))
        ux[kx // 2] = 0.0
        uy[ky // 2] = 0.0
    F = [rng.standard_normal((kx, ky)) + 1j * rng.standard_normal((kx, ky)) for _ in range(4)]
    F[2] = F[2] / fo.Z0
    F[3] = F[3] / fo.Z0
    ref = fo.radiated_power(F[0], F[1], F[2], F[3], ux, uy, d, d, wl, ng)
    got = power_from_projections(F, ux, uy, d, d, wl, ng)
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    fin = np.isfinite(ref)
    assert fin.any()
    assert np.abs(got - ref)[fin].max() <= 1e-12 * np.abs(ref[fin]).max()
