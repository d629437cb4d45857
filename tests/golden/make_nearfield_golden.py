"""Generate near-field golden vectors by running the UNMODIFIED reference
(/root/reference/{grating,lens_center,design_collimator,nearfield}.py) on the synthetic
lens library of tests/synth_lens.py.  Dev container only.

    python tests/golden/make_nearfield_golden.py
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import reference_loader  # noqa: E402
import synth_lens  # noqa: E402

ref = reference_loader.load()
G, LC, DC, NF = ref["grating"], ref["lens_center"], ref["design_collimator"], ref["nearfield"]
inf = float("inf")


def summarise(periph):
    return {("periph_" + k): np.asarray(v) for k, v in periph.items() if k != "gratingcollection_list"}


def run(name, spec, source, pol, wavelength=580e-9, x_pts=None, y_pts=None, big=False, **kw):
    collections, hgs = synth_lens.make_library(G, LC, spec)
    with contextlib.redirect_stdout(io.StringIO()):
        periph, center, r_switch = DC.make_design(collections, spec["source_distance"], spec["radius"], hgs)
        t0 = time.time()
        fn = NF.build_nearfield_big if big else NF.build_nearfield
        Ex, Ey, Hx, Hy, xp, yp, power, n_glass = fn(
            source_x=source[0], source_y=source[1], source_z=source[2], source_pol=pol, wavelength=wavelength,
            lens_periphery_summary=periph, lens_center_summary=center, hexgridset=hgs, x_pts=x_pts, y_pts=y_pts, **kw)
        dt = time.time() - t0
    out = dict(Ex=Ex, Ey=Ey, Hx=Hx, Hy=Hy, x_pts=xp, y_pts=yp, power=power, n_glass=n_glass,
               center=center, r_switch=r_switch, source=np.array(source, float), pol=pol, wavelength=wavelength)
    out.update(summarise(periph))
    np.savez_compressed(os.path.join(HERE, "nearfield_%s.npz" % name), **out)
    print(name, Ex.shape, "rings", len(periph["r_min_list"]), "cells", len(center), "power %.6e" % power,
          "max|Ex| %.3e" % np.abs(Ex).max(), "%.1fs" % dt)


spec = synth_lens.SMALL_LENS
f = spec["source_distance"]
run("small_x_onaxis", spec, (0.0, 0.0, -f), "x")
run("small_y_offaxis", spec, (0.7e-6, -0.4e-6, -f), "y")
run("small_z_onaxis", spec, (0.0, 0.0, -f), "z")
run("plane_x", synth_lens.PLANE_LENS, (0.0, 0.0, -inf), "x", dipole_moment=1.0)
run("plane_lens_y_point", synth_lens.PLANE_LENS, (0.3e-6, 0.2e-6, -synth_lens.PLANE_LENS["source_distance"]), "y")
# error behaviour: normal incidence is outside the SMALL_LENS tables -> ValueError (nearfield.py:294-295)
try:
    run("never", spec, (0.0, 0.0, -inf), "x")
    raise SystemExit("expected ValueError")
except ValueError as e:
    print("ValueError as expected:", e.args)
    np.savez(os.path.join(HERE, "nearfield_error_small_plane.npz"), message=e.args[0], value=e.args[1], bound=e.args[2])
# explicit, non-square, odd sample grid that over-fills the lens (points outside -> 0)
xs = np.linspace(-13e-6, 13e-6, 105)
ys = np.linspace(-9e-6, 12.5e-6, 87)
run("small_x_ragged", spec, (0.0, 0.0, -f), "x", x_pts=xs, y_pts=ys)
