"""Import the UNMODIFIED reference from /root/reference (dev container only).

Used solely by the ``make_*.py`` fixture generators in this directory; nothing
under ``tests/test_*.py``, ``bench.py`` or ``__graft_entry__.py`` imports it,
because /root/reference does not exist on the GPU box.

The reference needs five modules that are not installed here
(``numericalunits``, ``matplotlib``, ``dxfwrite``, ``ezdxf``, ``svgwrite``);
they are replaced by in-memory stubs.  ``numericalunits`` is pinned to SI, the
same constants as ``metalens_b200.units``.
"""
import math
import os
import sys
import types

REFERENCE_DIR = os.environ.get("METALENS_REFERENCE_DIR", "/root/reference")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Return dict of the reference modules on the hot path."""
    if not os.path.isdir(REFERENCE_DIR):
        raise RuntimeError("reference tree not found at %s" % REFERENCE_DIR)
    mu0 = 4e-7 * math.pi
    c0 = 299792458.0
    _stub("numericalunits", m=1.0, nm=1e-9, um=1e-6, C=1.0, V=1.0, c0=c0,
          mu0=mu0, eps0=1.0 / (mu0 * c0 ** 2), Z0=mu0 * c0,
          reset_units=lambda *a, **k: None)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.patches = _stub("matplotlib.patches")
    _stub("dxfwrite", DXFEngine=object)
    _stub("ezdxf")
    _stub("svgwrite")
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import nearfield_farfield, nearfield, grating, lens_center, design_collimator
    return dict(nearfield_farfield=nearfield_farfield, nearfield=nearfield,
                grating=grating, lens_center=lens_center,
                design_collimator=design_collimator)
