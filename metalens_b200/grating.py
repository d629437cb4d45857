"""Host-side mirror of the reference's ``grating.py`` for the hot path.

Keeps the ``Grating`` / ``GratingCollection`` call surface that ``build_nearfield`` and the
``design_collimator`` flow read: constructor arguments, ``.data`` rows
(schema grating.py:453-463), ``grating_period`` / ``lateral_period`` / ``n_glass``,
``get_angle_in_air``, ``get_one``, ``build_interpolators()`` with its ``.interpolators``
dict and ``.interpolator_bounds`` 6-tuple, ``repr`` round trip.

Out of scope (SURVEY component 3'): everything that drives the external S4 RCWA solver
(``characterize``, ``optimize``, ``vary_angle``, file writers, plotting).  ``characterize``
is kept as a named entry point that explains how to supply ``.data`` instead.
"""
import math

import numpy as np
from numpy import array  # noqa: F401  (repr() round trip evaluates `array(...)`)

from .tables import AmplitudeTable
from .units import nm, um, degree  # noqa: F401

pi = math.pi
inf = float('inf')

AMPS_GRATINGCOLLECTION = ('ampfy', 'ampfx')      # grating.py:1202-1203 builds only these two


def n_glass(wavelength_in_nm):
    """Substrate index table of grating.py:1274-1288 (ValueError for other wavelengths, Q6)."""
    table = {450: 1.466, 500: 1.462, 525: 1.461, 550: 1.46, 575: 1.459, 580: 1.459,
             600: 1.458, 625: 1.457, 650: 1.457}
    if wavelength_in_nm not in table:
        raise ValueError('bad wavelength' + repr(wavelength_in_nm))
    return table[wavelength_in_nm]


class Grating:
    """One periodic unit cell: ``grating_period`` x ``lateral_period`` (grating.py:95-141).

    Supply EITHER grating_period OR (angle_in_air and target_wavelength).  ``data`` is the
    list of characterisation rows ``{wavelength_in_nm, ux, uy, ox, oy, x_or_y, ampfy, ampfx,
    ampry, amprx}``.
    """

    def __init__(self, lateral_period, cyl_height, grating_period=None, target_wavelength=None,
                 angle_in_air=None, n_glass=0, n_tio2=0, xyrra_list_in_nm_deg=None, data=None):
        if grating_period is None:
            grating_period = target_wavelength / math.sin(angle_in_air)      # grating.py:127
        else:
            assert target_wavelength is None and angle_in_air is None        # grating.py:124
        self.grating_period = grating_period
        self.lateral_period = lateral_period
        self.cyl_height = cyl_height
        self.n_glass = n_glass
        self.n_tio2 = n_tio2
        self.grating_kx = 2 * pi / grating_period
        if xyrra_list_in_nm_deg is not None:
            xyrra = np.array(xyrra_list_in_nm_deg, dtype=float)
            xyrra[:, 0:4] *= nm
            xyrra[:, 4] *= degree
            self.xyrra_list = xyrra
        if data is not None:
            self.data = data

    @property
    def xyrra_list_in_nm_deg(self):
        out = self.xyrra_list.copy()
        out[:, 0:4] /= nm
        out[:, 4] /= degree
        return out

    def get_angle_in_air(self, target_wavelength):
        """Angle (from normal, in air) at which this cell sits in a lens designed for
        target_wavelength (grating.py:195-201)."""
        if self.grating_period < target_wavelength:
            raise ValueError('bad inputs!', target_wavelength / nm, self.grating_period / nm)
        return math.asin(target_wavelength / self.grating_period)

    def __repr__(self):
        if hasattr(self, 'xyrra_list'):
            xy = 'np.array(' + np.array2string(self.xyrra_list_in_nm_deg, separator=',').replace(' ', '').replace('\n', '') + ')'
        else:
            xy = 'None'
        return ('Grating(lateral_period=%r*nm, grating_period=%r*nm, cyl_height=%r*nm, n_glass=%r, n_tio2=%r, '
                'xyrra_list_in_nm_deg=%s, data=%s)'
                % (self.lateral_period / nm, self.grating_period / nm, self.cyl_height / nm, self.n_glass,
                   self.n_tio2, xy, repr(self.data) if hasattr(self, 'data') else 'None'))

    def copy(self):
        return eval(repr(self))                                              # grating.py:280-281

    def characterize(self, *args, **kwargs):
        raise NotImplementedError(
            "Grating.characterize() runs the external S4 RCWA solver (grating.py:364-478), which is outside "
            "this engine. Fill `.data` with rows from a saved repr() of a characterised grating instead.")


class GratingCollection:
    """Smoothly varying family of gratings for one band of angles (grating.py:920-958)."""

    def __init__(self, target_wavelength, lateral_period, lens_type='cyl', grating_list=None):
        assert lens_type in ('cyl', 'round')
        self.target_wavelength = target_wavelength
        self.lateral_period = lateral_period
        self.target_kvac = 2 * pi / target_wavelength
        self.lens_type = lens_type
        self.grating_list = [] if grating_list is None else grating_list
        if grating_list is not None:
            self.sort_grating_list()
            self.check_consistency()

    def check_consistency(self):
        """Same checks as grating.py:955-969."""
        gl = self.grating_list
        assert len({g.cyl_height for g in gl}) <= 1
        assert len({g.n_glass for g in gl}) <= 1
        assert len({g.n_tio2 for g in gl}) <= 1
        if self.lens_type == 'cyl':
            assert all(g.lateral_period == self.lateral_period for g in gl)
        else:
            ratio = [g.lateral_period / math.tan(g.get_angle_in_air(self.target_wavelength)) for g in gl]
            assert max(ratio) - min(ratio) < 1e-7 * max(ratio)

    def sort_grating_list(self):
        self.grating_list.sort(key=lambda g: g.grating_period)

    def add_one(self, new_grating):
        self.grating_list.append(new_grating)
        self.sort_grating_list()
        self.check_consistency()

    def get_innermost(self):
        return self.grating_list[-1]

    def get_outermost(self):
        return self.grating_list[0]

    def get_one(self, angle_in_air=None, grating_period=None, lateral_period=None):
        """Grating of this family at a given angle / period, pillar geometry linearly
        interpolated between the two neighbours (grating.py:981-1047)."""
        if grating_period is not None:
            assert angle_in_air is None and lateral_period is None
        elif angle_in_air is not None:
            assert lateral_period is None
            grating_period = self.target_wavelength / math.sin(angle_in_air)
        else:
            assert self.lens_type == 'round'
            grating_period = self.target_wavelength / math.sin(math.atan(lateral_period / self.lateral_period))
        if self.lens_type == 'cyl':
            lateral = self.lateral_period
        else:
            lateral = self.lateral_period * math.tan(math.asin(self.target_wavelength / grating_period))
        self.sort_grating_list()
        gl = self.grating_list
        periods = [g.grating_period for g in gl]
        has_geometry = all(hasattr(g, 'xyrra_list') for g in gl)
        if not has_geometry or grating_period < periods[0] * 0.99 or grating_period > periods[-1] * 1.01:
            xyrra = None
        elif grating_period >= periods[-1]:
            xyrra = gl[-1].xyrra_list_in_nm_deg
        elif grating_period <= periods[0]:
            xyrra = gl[0].xyrra_list_in_nm_deg
        elif grating_period in periods:
            xyrra = gl[periods.index(grating_period)].xyrra_list_in_nm_deg
        else:
            hi = next(j for j, p in enumerate(periods) if p > grating_period)
            w_hi = (grating_period - periods[hi - 1]) / (periods[hi] - periods[hi - 1])
            xyrra = (1 - w_hi) * gl[hi - 1].xyrra_list_in_nm_deg + w_hi * gl[hi].xyrra_list_in_nm_deg
        return Grating(lateral_period=lateral, cyl_height=gl[0].cyl_height, grating_period=grating_period,
                       n_glass=gl[0].n_glass, n_tio2=gl[0].n_tio2, xyrra_list_in_nm_deg=xyrra)

    def characterize(self, *args, **kwargs):
        raise NotImplementedError(
            "GratingCollection.characterize() fans out S4 processes (grating.py:1160-1184); outside this "
            "engine. Provide `.data` on every grating, then call build_interpolators().")

    def build_interpolators(self):
        """Dense tables over (ux, uy, grating_period) per (wavelength_nm, (ox,oy), pol, amp):
        zeros where an order was not reported (grating.py:1207-1214), period axis padded with
        copies of the edge planes at 0.99*min and 1.01*max (grating.py:1216-1226).  Sets
        ``self.interpolators`` and ``self.interpolator_bounds`` (grating.py:1229-1232).
        One pass over the rows instead of the reference's triple loop."""
        gl = self.grating_list
        ux_list = sorted({e['ux'] for g in gl for e in g.data})
        uy_list = sorted({e['uy'] for g in gl for e in g.data})
        gp_list = sorted({g.grating_period for g in gl})
        iu = {u: i for i, u in enumerate(ux_list)}
        iv = {u: i for i, u in enumerate(uy_list)}
        ig = {p: i for i, p in enumerate(gp_list)}
        wavelengths = {round(e['wavelength_in_nm']) for g in gl for e in g.data}
        orders = {(e['ox'], e['oy']) for g in gl for e in g.data}
        shape = (len(ux_list), len(uy_list), len(gp_list) + 2)
        dense = {(w, o, pol, amp): np.zeros(shape, dtype=complex)
                 for w in wavelengths for o in orders for pol in ('x', 'y') for amp in AMPS_GRATINGCOLLECTION}
        for g in gl:                                  # later gratings override equal periods, like the
            k = ig[g.grating_period] + 1              # reference's lookup dict (grating.py:1198-1199)
            for e in g.data:
                if e['x_or_y'] not in ('x', 'y'):
                    continue
                head = (round(e['wavelength_in_nm']), (e['ox'], e['oy']), e['x_or_y'])
                for amp in AMPS_GRATINGCOLLECTION:
                    dense[head + (amp,)][iu[e['ux']], iv[e['uy']], k] = e[amp]
        gp_ext = np.hstack((0.99 * min(gp_list), gp_list, 1.01 * max(gp_list)))
        self.interpolators = {}
        for key, arr in dense.items():
            arr[:, :, 0] = arr[:, :, 1]
            arr[:, :, -1] = arr[:, :, -2]
            self.interpolators[key] = AmplitudeTable((ux_list, uy_list, gp_ext), arr)
        self.interpolator_bounds = (min(ux_list), max(ux_list), min(uy_list), max(uy_list),
                                    min(gp_ext), max(gp_ext))

    def __repr__(self):
        return ('GratingCollection(target_wavelength=%r*nm, lateral_period=%r*nm, lens_type=%r, grating_list=%r)'
                % (self.target_wavelength / nm, self.lateral_period / nm, self.lens_type, self.grating_list))
