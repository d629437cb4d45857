"""CPU oracle for the near-field -> far-field transform.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain float64 numpy, the algorithm of the reference's
``nearfield_farfield.py`` (Taflove 1995 ch. 8 surface-equivalence transform).
It exists so that the CUDA path can be checked on a GPU box where
``/root/reference`` is not mounted.  Only ``tests/``, ``__graft_entry__.smoke()``
and the CPU legs of ``bench.py`` may import it; the product package
``metalens_b200`` never does.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned by *executing the unmodified
reference* in the dev container (``tests/golden/make_farfield_golden.py``) and
committing its outputs under ``tests/golden/``; ``tests/test_oracle_farfield.py``
checks this file against those fixtures to ~1e-13.

Every function cites the reference lines it follows.
"""
import math

import numpy as np

MU0 = 4e-7 * math.pi
C0 = 299792458.0
Z0 = MU0 * C0          # numericalunits.Z0 in SI; used at nearfield_farfield.py:183

REG_SIN = 1e-9         # "+ 1e-9" regulariser on sin(theta), nearfield_farfield.py:158-167
REG_UZ = 1e-5          # "+ 1e-5" regulariser on uz,         nearfield_farfield.py:185


def check_uniform_axis(pts, wavelength):
    """Grid validation of nearfield_farfield.py:26-30 (AssertionError on violation)."""
    pts = np.asarray(pts, dtype=float)
    steps = np.diff(pts)
    assert 0 < steps[0] < wavelength / 2
    assert steps.max() - steps.min() <= 1e-9 * np.abs(steps).max()


def fft_bin_direction_cosines(num, spacing, wavelength, n_glass):
    """Direction cosines (in glass) of the `num` FFT bins, un-shifted order.

    nearfield_farfield.py:35-39: u_i = i*(lambda/n)/(spacing*num); bins above half
    of the largest value are aliased down by (lambda/n)/spacing, so for even
    `num` the Nyquist bin is negative.
    """
    lam = wavelength / n_glass
    u = np.arange(num) * lam / (spacing * num)
    u[u > u.max() / 2] -= lam / spacing
    return u


def radiated_power(FEx, FEy, FHx, FHy, ux_list, uy_list, dxp, dyp, wavelength, n_glass):
    """Power per unit (ux,uy) from aperture-summed fields.

    Follows farfield_from_nearfield_helper, nearfield_farfield.py:135-189.
    ``FEx[i,j]`` must be sum_{x',y'} Ex e^{-ik(x' ux_i + y' uy_j)} (what
    fft2(fftshift(Ex)) is on the FFT-bin grid).  Evanescent bins come back NaN
    (:153-155).  The bin with ux == 0 and uy == 0 exactly uses the Cartesian
    components (:161-169).  The reference finds that bin with an indexing quirk
     that is only right for un-shifted FFT order; the *intent* is
    implemented here so any bin order works.
    """
    ux = np.asarray(ux_list, dtype=float).reshape(-1, 1)
    uy = np.asarray(uy_list, dtype=float).reshape(1, -1)
    area = dxp * dyp
    Nx = -FHy * area                      # J = n x H, (8.15), :135
    Ny = FHx * area                       # :136
    Lx = FEy * area                       # M = -n x E, :137
    Ly = -FEx * area                      # :138

    uz2 = 1 - ux ** 2 - uy ** 2           # :153
    uz = np.sqrt(np.where(uz2 < 0, np.nan, uz2))   # :154-155
    s = np.sqrt(ux ** 2 + uy ** 2)        # sin(theta), :156
    Nth = Nx * ux * uz / (s + REG_SIN) + Ny * uy * uz / (s + REG_SIN)    # :158
    Nph = -Nx * uy / (s + REG_SIN) + Ny * ux / (s + REG_SIN)            # :159
    Lth = Lx * ux * uz / (s + REG_SIN) + Ly * uy * uz / (s + REG_SIN)    # :166
    Lph = -Lx * uy / (s + REG_SIN) + Ly * ux / (s + REG_SIN)            # :167
    dc = (ux == 0) & (uy == 0)            # :161-165, :168-169 (intent)
    if dc.any():
        Nth[dc], Nph[dc] = Nx[dc], Ny[dc]
        Lth[dc], Lph[dc] = Lx[dc], Ly[dc]

    Z = Z0 / n_glass                      # :183
    pref = (2 * math.pi * n_glass / wavelength) ** 2 / (32 * math.pi ** 2 * Z)
    P = pref * (np.abs(Lph + Z * Nth) ** 2 + np.abs(Lth - Z * Nph) ** 2) / (uz + REG_UZ)  # :184-185
    P *= 2                                # "mystery factor", :189
    return P


def farfield_from_fft(fftEx, fftEy, fftHx, fftHy, xp_list, yp_list, wavelength, n_glass):
    """Same contract as the reference's farfield_from_nearfield (nearfield_farfield.py:14-75).

    Inputs are fft2(fftshift(field)).  Returns
    (P fftshifted (M,M), total_P, ux (M,1), uy (1,M), dux, duy).
    The reference's RAM chunk loop (:45-66) is bookkeeping and is not restated.
    """
    dxp = xp_list[1] - xp_list[0]
    dyp = yp_list[1] - yp_list[0]
    nx, ny = len(xp_list), len(yp_list)
    assert fftEx.shape == fftEy.shape == fftHx.shape == fftHy.shape == (nx, ny)   # :26
    check_uniform_axis(xp_list, wavelength)
    check_uniform_axis(yp_list, wavelength)
    ux = fft_bin_direction_cosines(nx, dxp, wavelength, n_glass)
    uy = fft_bin_direction_cosines(ny, dyp, wavelength, n_glass)
    P = radiated_power(fftEx, fftEy, fftHx, fftHy, ux, uy, dxp, dyp, wavelength, n_glass)
    P = np.fft.fftshift(P)                # :68
    ux = np.fft.fftshift(ux)              # :69
    uy = np.fft.fftshift(uy)              # :70
    dux = ux[1] - ux[0]                   # :71
    duy = uy[1] - uy[0]                   # :72
    total_P = (P * dux * duy)[np.isfinite(P)].sum()    # :74
    return P, total_P, ux.reshape(-1, 1), uy.reshape(1, -1), dux, duy


def farfield_reference_path(Ex, Ey, Hx, Hy, xp_list, yp_list, wavelength, n_glass):
    """The whole reference CPU path: 4x fft2(fftshift(.)) (caller side,
    nearfield_farfield.py:18-20) followed by farfield_from_fft.  This is what the
    CPU baseline in bench.py times."""
    f = [np.fft.fft2(np.fft.fftshift(np.asarray(a, dtype=complex))) for a in (Ex, Ey, Hx, Hy)]
    return farfield_from_fft(f[0], f[1], f[2], f[3], xp_list, yp_list, wavelength, n_glass)


def farfield_reference_path_threads(Ex, Ey, Hx, Hy, xp_list, yp_list, wavelength, n_glass, workers,
                                    points_at_a_time=1e7):
    """farfield_reference_path() with its independent pieces spread over `workers` threads -- the
    "all host cores" flavour of the CPU baseline.  Two things are independent in the reference itself:
    the four caller-side fft2(fftshift(.)) calls (nearfield_farfield.py:18-20) and the uy-chunk loop of
    farfield_from_nearfield (:45-66, chunks of `points_at_a_time`/num_x columns; here the chunk width is
    additionally capped so that every worker has work).  numpy's pocketfft and ufunc loops release the
    GIL, so plain threads scale.  Elementwise arithmetic is chunk-invariant and total_P is summed over the
    assembled map exactly as at :74, so the result is bit-identical to farfield_reference_path()."""
    from concurrent.futures import ThreadPoolExecutor
    dxp = xp_list[1] - xp_list[0]
    dyp = yp_list[1] - yp_list[0]
    nx, ny = len(xp_list), len(yp_list)
    assert Ex.shape == Ey.shape == Hx.shape == Hy.shape == (nx, ny)               # :26
    check_uniform_axis(xp_list, wavelength)
    check_uniform_axis(yp_list, wavelength)
    ux = fft_bin_direction_cosines(nx, dxp, wavelength, n_glass)
    uy = fft_bin_direction_cosines(ny, dyp, wavelength, n_glass)
    cols = max(1, min(int(points_at_a_time / nx), -(-ny // max(1, workers))))     # :46, capped
    starts = list(range(0, ny, cols))                                             # :47-49
    with ThreadPoolExecutor(max(1, workers)) as pool:
        f = list(pool.map(lambda a: np.fft.fft2(np.fft.fftshift(np.asarray(a, dtype=complex))), (Ex, Ey, Hx, Hy)))

        def chunk(j0):                                                            # body of the loop, :50-66
            sl = slice(j0, min(ny, j0 + cols))
            return radiated_power(f[0][:, sl], f[1][:, sl], f[2][:, sl], f[3][:, sl], ux, uy[sl], dxp, dyp,
                                  wavelength, n_glass)
        P = np.concatenate(list(pool.map(chunk, starts)), axis=1)
    P = np.fft.fftshift(P)                # :68
    ux = np.fft.fftshift(ux)              # :69
    uy = np.fft.fftshift(uy)              # :70
    dux = ux[1] - ux[0]                   # :71
    duy = uy[1] - uy[0]                   # :72
    total_P = (P * dux * duy)[np.isfinite(P)].sum()    # :74
    return P, total_P, ux.reshape(-1, 1), uy.reshape(1, -1), dux, duy


def fftshift_origin_index(num):
    """Index of the aperture sample that fftshift() moves to position 0, i.e. the
    phase origin x'=0 of the reference transform (nearfield_farfield.py:106-110;
    SURVEY Q4).  numpy's fftshift rolls by num//2, so this is num - num//2."""
    return num - num // 2


def aperture_sum_dense(J, x_rel, y_rel, ux_list, uy_list, wavelength, n_glass):
    """Direct separable sum  F[i,j] = sum_{m1,m2} J[m1,m2] e^{-ik(x'_{m1} ux_i + y'_{m2} uy_j)}
    in float64 (derivation: nearfield_farfield.py:97-120; k = 2 pi n / lambda).
    x_rel, y_rel are sample coordinates relative to the phase origin."""
    k = 2 * math.pi * n_glass / wavelength
    Ax = np.exp(-1j * k * np.outer(np.asarray(ux_list, float), np.asarray(x_rel, float)))   # (K, M)
    Ay = np.exp(-1j * k * np.outer(np.asarray(y_rel, float), np.asarray(uy_list, float)))   # (M, K)
    return Ax @ np.asarray(J, dtype=complex) @ Ay


def farfield_dense(Ex, Ey, Hx, Hy, dxp, dyp, ux_list, uy_list, wavelength, n_glass,
                   origin_x=None, origin_y=None):
    """Far-field power on an ARBITRARY (ux_list x uy_list) grid by direct summation
    + radiated_power().  With ux/uy taken from the FFT-bin grid this equals the
    reference output at those bins (SURVEY KAT-4)."""
    nx, ny = Ex.shape
    ox = fftshift_origin_index(nx) if origin_x is None else origin_x
    oy = fftshift_origin_index(ny) if origin_y is None else origin_y
    x_rel = (np.arange(nx) - ox) * dxp
    y_rel = (np.arange(ny) - oy) * dyp
    F = [aperture_sum_dense(a, x_rel, y_rel, ux_list, uy_list, wavelength, n_glass)
         for a in (Ex, Ey, Hx, Hy)]
    P = radiated_power(F[0], F[1], F[2], F[3], ux_list, uy_list, dxp, dyp, wavelength, n_glass)
    return P, F
