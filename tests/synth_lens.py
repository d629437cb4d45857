"""Deterministic synthetic metalens library (no S4 needed), SURVEY.md section 8d cfg4.

`characterize()` of the reference needs the external S4 RCWA binary, which is not
available; the amplitude tables (`Grating.data`, row schema grating.py:453-463) are
therefore synthetic closed-form functions of (order, incidence direction, period).
The same generator is run against the reference's classes (golden generation, dev
container only) and against `metalens_b200`'s classes (tests, bench), so both sides
see bit-identical tables.
"""
import math

import numpy as np

nm = 1e-9
um = 1e-6
degree = math.pi / 180

SEP = 320 * nm              # design_collimator.py:34 pitch
CYL_HEIGHT = 550 * nm       # design_collimator.py:36
N_GLASS = 1.459             # grating.n_glass(580)
N_TIO2 = 2.372
WAVELENGTH = 580 * nm       # design_collimator.py:51 design wavelength
MAX_ORDER = 3


def synthetic_amp(kind, pol, ox, oy, ux, uy, period_nm, salt):
    """Smooth, deterministic complex amplitude with |a| < 1."""
    k = {"ampfy": 0.0, "ampfx": 1.0, "ampry": 2.0, "amprx": 3.0}[kind]
    p = 0.0 if pol == "x" else 1.0
    mag = 0.45 + 0.3 * math.sin(0.9 * ox - 1.7 * oy + 2.1 * ux - 1.3 * uy + 0.004 * period_nm + 0.8 * k + 1.9 * p + salt)
    mag *= 1.0 / (1.0 + 0.35 * (ox * ox + oy * oy))
    if kind == "ampfx" and pol == "x" and ox == 0 and oy == 0:
        mag += 0.4
    ph = 1.1 * ux + 0.6 * uy + 0.5 * ox - 0.8 * oy + 0.003 * period_nm + 0.7 * k - 0.4 * p + 0.37 * salt
    return complex(mag * math.cos(ph), mag * math.sin(ph))


def table_rows(wavelength_nm, ux_list, uy_list, grating_period, lateral_period, salt):
    """`Grating.data` rows: for every (ux,uy) node, both incident polarisations and every
    order |ox|,|oy| <= MAX_ORDER that propagates in air at that node (grating.lua:416-423)."""
    rows = []
    lam = wavelength_nm * nm
    for ux in ux_list:
        for uy in uy_list:
            for ox in range(-MAX_ORDER, MAX_ORDER + 1):
                for oy in range(-MAX_ORDER, MAX_ORDER + 1):
                    sx = ux + ox * lam / grating_period
                    sy = uy + oy * lam / lateral_period
                    if sx * sx + sy * sy > 1.0:
                        continue
                    for pol in ("x", "y"):
                        row = {"wavelength_in_nm": float(wavelength_nm), "ux": float(ux), "uy": float(uy),
                               "ox": ox, "oy": oy, "x_or_y": pol}
                        for kind in ("ampfy", "ampfx", "ampry", "amprx"):
                            row[kind] = synthetic_amp(kind, pol, ox, oy, ux, uy, grating_period / nm, salt)
                        rows.append(row)
    return rows


def make_collection(grating_mod, angle_lo, angle_hi, lateral_over_tan, n_gratings=5, u_steps=5,
                    wavelength_nm=580, salt=0.0, n_glass=N_GLASS):
    """Round-lens GratingCollection covering incidence angles [angle_lo, angle_hi] (radians)."""
    # a little wider than the nominal range so every ring period is inside the table (the
    # reference pads the period axis by only 1 %, grating.py:1219-1226)
    angles = np.linspace(angle_lo - 1.5 * degree, angle_hi + 1.5 * degree, n_gratings)
    ux_min = max(-0.99, math.sin(angle_lo) - 0.25)        # grating.py:1167-1170
    ux_max = min(0.99, math.sin(angle_hi) + 0.25)
    ux_list = np.linspace(ux_min, ux_max, u_steps)
    uy_list = np.linspace(-0.2, 0.2, u_steps)             # grating.py:1171-1172
    gratings = []
    for a in angles:
        gp = WAVELENGTH / math.sin(a)
        lat = lateral_over_tan * math.tan(a)
        g = grating_mod.Grating(lateral_period=lat, cyl_height=CYL_HEIGHT, grating_period=gp,
                                n_glass=n_glass, n_tio2=N_TIO2)
        g.data = table_rows(wavelength_nm, ux_list, uy_list, gp, lat, salt)
        gratings.append(g)
    return grating_mod.GratingCollection(target_wavelength=WAVELENGTH, lateral_period=lateral_over_tan,
                                         lens_type="round", grating_list=gratings)


def make_hexgridset(lens_center_mod, grating_mod, num_entries=20, u_steps=5, wavelength_nm=580, n_glass=N_GLASS):
    ux_list = np.linspace(-0.499, 0.501, u_steps)         # lens_center.py:92-94
    uy_list = np.linspace(-0.499, 0.501, u_steps)
    gratings = []
    for k in range(num_entries):
        g = grating_mod.Grating(grating_period=SEP * 3 ** 0.5, lateral_period=SEP, n_glass=n_glass,
                                n_tio2=N_TIO2, cyl_height=CYL_HEIGHT)
        g.data = table_rows(wavelength_nm, ux_list, uy_list, SEP * 3 ** 0.5, SEP, salt=0.21 * k)
        gratings.append(g)
    x_amp = [0.9 * complex(math.cos(2 * math.pi * k / num_entries), math.sin(2 * math.pi * k / num_entries))
             for k in range(num_entries)]
    return lens_center_mod.HexGridSet(sep=SEP, cyl_height=CYL_HEIGHT, n_glass=n_glass, n_tio2=N_TIO2,
                                      grating_list=gratings, x_amp_list=x_amp)


# (angle_lo, angle_hi, lateral_period/tan(angle), salt) of the synthetic periphery
SMALL_LENS = dict(bands=[(15.0, 25.0, 1000 * nm, 0.3), (25.0, 40.0, 650 * nm, 1.1)],
                  source_distance=14.3 * um, radius=11.5 * um)

# bands starting below asin(0.25) so that normal incidence (ux = 0) is inside the tables
PLANE_LENS = dict(bands=[(5.0, 9.0, 3600 * nm, 0.5), (9.0, 13.0, 2200 * nm, 1.7)],
                  source_distance=60.0 * um, radius=13.5 * um)


def make_library(grating_mod, lens_center_mod, spec=SMALL_LENS, wavelength_nm=580, n_glass=N_GLASS):
    """Returns (collections list for make_design, hexgridset), interpolators built."""
    collections = []
    for lo, hi, lot, salt in spec["bands"]:
        gc = make_collection(grating_mod, lo * degree, hi * degree, lot, salt=salt, wavelength_nm=wavelength_nm,
                             n_glass=n_glass)
        gc.build_interpolators()
        collections.append([(lo * degree, hi * degree), gc])
    hgs = make_hexgridset(lens_center_mod, grating_mod, wavelength_nm=wavelength_nm, n_glass=n_glass)
    hgs.build_interpolators()
    return collections, hgs

# SURVEY section 8 probe-sized lens: default grid 675 x 675 (= good_fft_number), ~20 rings, ~1.2e5 hex cells
MID_LENS = dict(bands=[(15.0, 25.0, 1000 * nm, 0.3)], source_distance=216.5 * um, radius=88.9 * um)
