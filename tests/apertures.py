"""Seeded synthetic aperture fields shared by the tests, bench.py and the golden
generators.  Pure numpy, no reference imports."""
import math

import numpy as np

MU0 = 4e-7 * math.pi
C0 = 299792458.0
Z0 = MU0 * C0

N_GLASS = {450: 1.466, 532: 1.4607, 580: 1.459, 635: 1.457}


def grid(M, wavelength, My=None):
    """Aperture sample coordinates with spacing lambda/2.2 (nearfield.py:96)."""
    d = wavelength / 2.2
    My = M if My is None else My
    x = (np.arange(M) - (M - 1) / 2.0) * d
    y = (np.arange(My) - (My - 1) / 2.0) * d
    return x, y


def uniform(M, wavelength, n_glass):
    """KAT-1: Ex=1, Hy=n/Z0 over the full grid."""
    x, y = grid(M, wavelength)
    Ex = np.ones((M, M), complex)
    Hy = Ex * n_glass / Z0
    z = np.zeros_like(Ex)
    return Ex, z, z.copy(), Hy, x, y


def disc(M, wavelength, n_glass, radius_samples=40):
    """KAT-2: uniform field inside a disc of radius 40 samples."""
    x, y = grid(M, wavelength)
    X, Y = np.meshgrid(x, y, indexing='ij')
    d = wavelength / 2.2
    Ex = ((X ** 2 + Y ** 2) <= (radius_samples * d) ** 2).astype(complex)
    Hy = Ex * n_glass / Z0
    z = np.zeros_like(Ex)
    return Ex, z, z.copy(), Hy, x, y


def tilted_te(M, wavelength, n_glass, angle_deg=20.0):
    """KAT-3: TE plane wave in glass tilted by angle_deg in the x-z plane
    (E along y).  E = -y_hat e^{ik sin(t) x}... sign chosen so power flows to +z."""
    x, y = grid(M, wavelength)
    X, _ = np.meshgrid(x, y, indexing='ij')
    k = 2 * math.pi * n_glass / wavelength
    t = math.radians(angle_deg)
    ph = np.exp(1j * k * math.sin(t) * X)
    Ey = ph.copy()
    Hx = -ph * n_glass / Z0 * math.cos(t)
    z = np.zeros_like(ph)
    return z, Ey, Hx, z.copy(), x, y


def gaussian_random(M, seed, wavelength, My=None, dtype=np.complex64):
    """cfg1(b): four i.i.d. N(0,1)+iN(0,1) fields, cast to complex64."""
    rng = np.random.default_rng(seed)
    My = M if My is None else My
    out = []
    for _ in range(4):
        a = rng.standard_normal((M, My)) + 1j * rng.standard_normal((M, My))
        out.append(a.astype(dtype))
    x, y = grid(M, wavelength, My)
    return out[0], out[1], out[2], out[3], x, y


def focusing_lens(M, seed, wavelength, n_glass, na=0.6, rotate=False, noise=0.01,
                  dtype=np.complex64):
    """cfg2/cfg3 generator: unit-amplitude converging wave exp(-ik_vac(sqrt(f^2+r^2)-f))
    inside a disc, H = n z_hat x E / Z0, plus 1 % complex Gaussian noise; `rotate`
    turns E,H by 90 degrees (second polarisation)."""
    rng = np.random.default_rng(seed)
    x, y = grid(M, wavelength)
    X, Y = np.meshgrid(x.astype(np.float64), y.astype(np.float64), indexing='ij', sparse=True)
    R = x[-1]
    f = R / math.tan(math.asin(na))
    r2 = X ** 2 + Y ** 2
    kv = 2 * math.pi / wavelength
    E = np.exp(-1j * kv * (np.sqrt(f * f + r2) - f)) * (r2 <= R * R)
    E = E + noise * (rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M)))
    E = E.astype(dtype)
    Hs = (E * (n_glass / Z0)).astype(dtype)
    z = np.zeros_like(E)
    if not rotate:      # E along x, H along y
        return E, z, z.copy(), Hs, x, y
    return z, E, (-Hs).astype(dtype), z.copy(), x, y
