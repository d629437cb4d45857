"""Generate far-field golden vectors by running the UNMODIFIED reference
(/root/reference/nearfield_farfield.py) in the dev container.

    python tests/golden/make_farfield_golden.py

Writes tests/golden/farfield_*.npz.  Inputs are NOT stored: tests regenerate
them from tests/apertures.py with the seeds recorded here.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import reference_loader  # noqa: E402
import apertures  # noqa: E402

ref = reference_loader.load()["nearfield_farfield"]
WL = 532e-9
NG = apertures.N_GLASS[532]


def run_reference(Ex, Ey, Hx, Hy, x, y, wl, ng):
    f = [np.fft.fft2(np.fft.fftshift(a.astype(complex))) for a in (Ex, Ey, Hx, Hy)]
    with contextlib.redirect_stdout(io.StringIO()):
        P, total, ux, uy, dux, duy = ref.farfield_from_nearfield(
            f[0], f[1], f[2], f[3], list(x), list(y), wl, ng)
    return dict(P=P, total_P=total, ux=ux.ravel(), uy=uy.ravel(), dux=dux, duy=duy,
                wavelength=wl, n_glass=ng)


def incident_power(Ex, Ey, Hx, Hy, x, y):
    return float(((Ex * np.conj(Hy) - Ey * np.conj(Hx)).real.sum()) * (x[1] - x[0]) * (y[1] - y[0]))


cases = {
    "kat1_uniform": apertures.uniform(128, WL, NG),
    "kat2_disc": apertures.disc(128, WL, NG),
    "kat3_tilted": apertures.tilted_te(128, WL, NG),
    "rand128_seed0": apertures.gaussian_random(128, 0, WL),
    "rand_48x40_seed5": apertures.gaussian_random(48, 5, WL, My=40),
    "rand_45x27_seed6": apertures.gaussian_random(45, 6, WL, My=27),
    "lens256_seed1": apertures.focusing_lens(256, 1, WL, NG),
    "lens256_seed1_rot": apertures.focusing_lens(256, 1, WL, NG, rotate=True),
}
for name, (Ex, Ey, Hx, Hy, x, y) in cases.items():
    out = run_reference(Ex, Ey, Hx, Hy, x, y, WL, NG)
    out["P_in"] = incident_power(Ex, Ey, Hx, Hy, x, y)
    np.savez_compressed(os.path.join(HERE, "farfield_%s.npz" % name), **out)
    print(name, "total_P/P_in =", out["total_P"] / out["P_in"] if out["P_in"] else None,
          "nan bins:", int(np.isnan(out["P"]).sum()))
