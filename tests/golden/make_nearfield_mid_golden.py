"""Mid-size near-field golden (SURVEY section 8 probe lens: 675 x 675 default grid, ~20 rings, ~1.2e5 hex cells)
from the UNMODIFIED reference, off-axis z- and y-dipoles.  Dev container only (~10 s per case).

The full 675^2 x 4 complex128 fields would be 29 MB per case; the fixture keeps the reference's values at
  * a seeded random subset of 24000 samples, and
  * EVERY sample that is a candidate for the two measure-zero classes where the reference's answer hinges on the
    last bit of a library call: centre samples whose two nearest hex cells are exactly
    equidistant (scipy cKDTree returns whichever its traversal meets first, nearfield.py:363-364) and periphery samples
    within 1e-9 of a boundary between two grating copies (round(phi / angle_per_grating), nearfield.py:167-169),
together with flags saying which is which, the incident power, and the reference's own make_design output.

    python tests/golden/make_nearfield_mid_golden.py
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
from scipy.spatial import cKDTree

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import reference_loader  # noqa: E402
import synth_lens  # noqa: E402

ref = reference_loader.load()
G, LC, DC, NF = ref["grating"], ref["lens_center"], ref["design_collimator"], ref["nearfield"]


def run(name, spec, source, pol, wavelength=580e-9, n_random=24000, seed=7, n_grid=None):
    collections, hgs = synth_lens.make_library(G, LC, spec)
    with contextlib.redirect_stdout(io.StringIO()):
        periph, center, r_switch = DC.make_design(collections, spec["source_distance"], spec["radius"], hgs)
        t0 = time.time()
        grid = None if n_grid is None else np.linspace(-periph["r_max_list"][-1], periph["r_max_list"][-1], n_grid)
        Ex, Ey, Hx, Hy, xp, yp, power, n_glass = NF.build_nearfield(
            source_x=source[0], source_y=source[1], source_z=source[2], source_pol=pol, wavelength=wavelength,
            lens_periphery_summary=periph, lens_center_summary=center, hexgridset=hgs, x_pts=grid, y_pts=grid)
        dt = time.time() - t0
    nx, ny = Ex.shape
    X, Y = np.meshgrid(xp, yp, indexing="ij")
    r = np.sqrt(X ** 2 + Y ** 2)
    # class 1: exact nearest-cell ties inside the centre region
    in_center = r < periph["r_min_list"][0]
    pts = np.stack((X[in_center], Y[in_center]), axis=1)
    d2, idx2 = cKDTree(center[:, 0:2]).query(pts, k=2)
    # exact comparison of the squared distances as the kernel forms them
    c0, c1 = center[idx2[:, 0], 0:2], center[idx2[:, 1], 0:2]
    q0 = (c0[:, 0] - pts[:, 0]) ** 2 + (c0[:, 1] - pts[:, 1]) ** 2
    q1 = (c1[:, 0] - pts[:, 0]) ** 2 + (c1[:, 1] - pts[:, 1]) ** 2
    tie = np.zeros((nx, ny), bool)
    tie[in_center] = (q0 == q1) | (d2[:, 0] == d2[:, 1])
    # class 2: periphery samples (numerically) on a boundary between two grating copies
    ring = np.searchsorted(np.hstack((periph["r_min_list"], periph["r_max_list"][-1])), r) - 1
    in_ring = (ring >= 0) & (ring < len(periph["r_min_list"]))
    apg = 2 * np.pi / np.asarray(periph["num_around_circle_list"], float)[np.clip(ring, 0, len(periph["r_min_list"]) - 1)]
    q = np.arctan2(Y, X) / apg
    wedge = in_ring & (np.abs(np.abs(q - np.round(q)) - 0.5) < 1e-9)
    rng = np.random.default_rng(seed)
    pick = np.zeros(nx * ny, bool)
    pick[rng.choice(nx * ny, size=n_random, replace=False)] = True
    pick |= tie.ravel() | wedge.ravel()
    flat = np.flatnonzero(pick)
    out = dict(index=flat.astype(np.int32), tie=tie.ravel()[flat], wedge=wedge.ravel()[flat],
               Ex=Ex.ravel()[flat], Ey=Ey.ravel()[flat], Hx=Hx.ravel()[flat], Hy=Hy.ravel()[flat],
               shape=np.array([nx, ny]), x_pts=xp, y_pts=yp, power=power, n_glass=n_glass,
               scale_E=max(np.abs(Ex).max(), np.abs(Ey).max()), scale_H=max(np.abs(Hx).max(), np.abs(Hy).max()),
               sum_abs2=np.array([np.sum(np.abs(a) ** 2) for a in (Ex, Ey, Hx, Hy)]),
               center=center, r_switch=r_switch, source=np.array(source, float), pol=pol, wavelength=wavelength)
    out.update({("periph_" + k): np.asarray(v) for k, v in periph.items() if k != "gratingcollection_list"})
    path = os.path.join(HERE, "nearfield_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, Ex.shape, "rings", len(periph["r_min_list"]), "cells", len(center), "ties", int(tie.sum()),
          "wedge", int(wedge.sum()), "stored", flat.size, "%.1fs" % dt, "%.2f MB" % (os.path.getsize(path) / 1e6))


spec = synth_lens.MID_LENS
f = spec["source_distance"]
run("mid_z_offaxis", spec, (3.1e-6, -1.7e-6, -f), "z")
# odd sample count: the rows x = 0 and y = 0 exist, where hex cells are exactly equidistant (symmetric lattice)
run("mid_y_onaxis", spec, (0.0, 0.0, -f), "y", n_grid=675)
