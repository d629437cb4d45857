"""Host-side mirror of the reference's ``lens_center.py``: the HexGridSet used for the
centre of the lens (hexagonal lattice of nano-pillars, one of ``num_entries`` diameters per
cell).  Keeps what ``design_center`` and ``build_nearfield`` read:
``grating_list``, ``x_amp_list``, ``pick_from_phase()``, ``build_interpolators()``.
``characterize()`` (S4-bound, lens_center.py:80-156) is out of scope.
"""
import math

import numpy as np

from . import grating
from .tables import AmplitudeTable
from .units import nm

pi = math.pi
AMPS_HEXGRIDSET = ('ampfy', 'ampfx', 'ampry', 'amprx')       # lens_center.py:206


class HexGridSet:
    """Set of hexagonal-lattice unit cells (lens_center.py:25-57).  ``sep`` is the
    nearest-neighbour pillar separation; every entry of ``grating_list`` is a Grating whose
    rectangular cell is (sep*sqrt(3)) x sep with two pillars."""

    def __init__(self, sep, cyl_height, n_glass=0, n_tio2=0, grating_list=None, x_amp_list=None, num_entries=20):
        self.sep = sep
        self.nnn_sep = sep * 3 ** 0.5
        self.cyl_height = cyl_height
        self.n_glass = n_glass
        self.n_tio2 = n_tio2
        if grating_list is None:
            grating_list = []
            for diam in np.linspace(100.01 * nm, sep - 100.01 * nm, num=num_entries):     # lens_center.py:42
                r = diam / 2 / nm
                cell = np.array([[0, 0, r, r, 0], [self.nnn_sep / 2 / nm, sep / 2 / nm, r, r, 0]])
                grating_list.append(grating.Grating(grating_period=self.nnn_sep, lateral_period=sep,
                                                    n_glass=n_glass, n_tio2=n_tio2, cyl_height=cyl_height,
                                                    xyrra_list_in_nm_deg=cell))
        self.grating_list = grating_list
        if x_amp_list is not None:
            self.x_amp_list = np.array(x_amp_list)

    def __repr__(self):
        amps = (np.array2string(self.x_amp_list, separator=',').replace(' ', '').replace('\n', '')
                if hasattr(self, 'x_amp_list') else 'None')
        return ('HexGridSet(sep=%r*nm, cyl_height=%r*nm, n_glass=%r, n_tio2=%r, grating_list= %r, x_amp_list=%s)'
                % (self.sep / nm, self.cyl_height / nm, self.n_glass, self.n_tio2, self.grating_list, amps))

    def characterize(self, *args, **kwargs):
        raise NotImplementedError(
            "HexGridSet.characterize() runs the external S4 solver (lens_center.py:80-156); outside this engine. "
            "Provide `.data` on every grating and `x_amp_list`, then call build_interpolators().")

    def pick_from_phase(self, target_phase):
        """Index of the cell whose normal-incidence x amplitude best matches target_phase
        (lens_center.py:175-186)."""
        if not hasattr(self, 'x_amp_list'):
            raise ValueError('Need to run characterize() first')
        return int(np.argmax((self.x_amp_list * np.exp(-1j * target_phase)).imag))

    def build_interpolators(self):
        """Tables over (ux, uy, grating index) for the four amplitudes (lens_center.py:188-226);
        zeros where a row is missing; ``interpolator_bounds`` = (ux min/max, uy min/max, 0, n-1)."""
        if not hasattr(self, 'x_amp_list'):
            raise ValueError('Need to run characterize() first')                      # lens_center.py:196-197
        gl = self.grating_list
        ux_list = sorted({e['ux'] for g in gl for e in g.data})
        uy_list = sorted({e['uy'] for g in gl for e in g.data})
        index_list = np.arange(len(gl))
        iu = {u: i for i, u in enumerate(ux_list)}
        iv = {u: i for i, u in enumerate(uy_list)}
        wavelengths = {round(e['wavelength_in_nm']) for g in gl for e in g.data}
        orders = {(e['ox'], e['oy']) for g in gl for e in g.data}
        shape = (len(ux_list), len(uy_list), len(gl))
        dense = {(w, o, pol, amp): np.zeros(shape, dtype=complex)
                 for w in wavelengths for o in orders for pol in ('x', 'y') for amp in AMPS_HEXGRIDSET}
        filled = set()
        for k, g in enumerate(gl):
            for e in g.data:
                # the reference matches rows with `e['wavelength_in_nm'] == wavelength_in_nm` (:213):
                # only rows whose wavelength is a whole number of nm are ever picked up
                if e['wavelength_in_nm'] != round(e['wavelength_in_nm']) or e['x_or_y'] not in ('x', 'y'):
                    continue
                head = (round(e['wavelength_in_nm']), (e['ox'], e['oy']), e['x_or_y'])
                cell = head + (iu[e['ux']], iv[e['uy']], k)
                assert cell not in filled, "duplicate characterisation row"           # lens_center.py:218
                filled.add(cell)
                for amp in AMPS_HEXGRIDSET:
                    dense[head + (amp,)][iu[e['ux']], iv[e['uy']], k] = e[amp]
        self.interpolators = {key: AmplitudeTable((ux_list, uy_list, index_list), arr) for key, arr in dense.items()}
        self.interpolator_bounds = (min(ux_list), max(ux_list), min(uy_list), max(uy_list),
                                    min(index_list), max(index_list))
