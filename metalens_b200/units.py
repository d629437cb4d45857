"""SI unit scale factors used by the host side.

The reference expresses every quantity through the ``numericalunits`` package
(`nearfield.py:14-15`, `nearfield_farfield.py:10`), which randomises its unit
scale per process unless ``reset_units('SI')`` is called.  This package fixes SI
once: lengths are metres, fields V/m and A/m, so the numbers handed to the CUDA
kernels never depend on a per-process random scale.
"""
import math

m = 1.0
um = 1e-6
nm = 1e-9
C = 1.0
V = 1.0
c0 = 299792458.0
mu0 = 4e-7 * math.pi
eps0 = 1.0 / (mu0 * c0 ** 2)
Z0 = mu0 * c0          # impedance of free space used at nearfield_farfield.py:183
degree = math.pi / 180.0
