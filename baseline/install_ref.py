"""Install the UNMODIFIED reference into baseline/_ref/ (git-ignored; it still travels to the GPU box with
gpurun, so `bench.py --impl reference` can time the reference's own code there).

The reference has no setup.py / pyproject.toml, so "install" = copy its Python modules as they
are.  Five modules it imports are not in this image (numericalunits, matplotlib, dxfwrite, ezdxf, svgwrite) and
there is no network: tiny stand-ins are generated under baseline/_ref/_stubs/ (numericalunits pinned to SI, the
constants of metalens_b200/units.py; the others are only imported at module level, never called on this path).
"""
import math
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUB_DIR = os.path.join(REF_DIR, "_stubs")
MODULES = ("nearfield_farfield.py", "nearfield.py", "grating.py", "lens_center.py", "design_collimator.py",
           "refractive_index.py")

_NUMERICALUNITS = '''"""SI stand-in for the `numericalunits` package (not installed; no network)."""
import math
m = 1.0
nm = 1e-9
um = 1e-6
C = 1.0
V = 1.0
c0 = 299792458.0
mu0 = 4e-7 * math.pi
eps0 = 1.0 / (mu0 * c0 ** 2)
Z0 = mu0 * c0


def reset_units(*a, **k):
    pass
'''


def install(source="/root/reference"):
    """Copy the reference modules (only when the source tree is present: dev container) and write the stubs."""
    if not os.path.isdir(source):
        return os.path.isdir(REF_DIR)
    os.makedirs(STUB_DIR, exist_ok=True)
    for name in MODULES:
        src = os.path.join(source, name)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(REF_DIR, name))
    with open(os.path.join(STUB_DIR, "numericalunits.py"), "w") as f:
        f.write(_NUMERICALUNITS)
    os.makedirs(os.path.join(STUB_DIR, "matplotlib"), exist_ok=True)
    for name in ("__init__.py", "pyplot.py", "patches.py"):
        open(os.path.join(STUB_DIR, "matplotlib", name), "w").close()
    with open(os.path.join(STUB_DIR, "dxfwrite.py"), "w") as f:
        f.write("DXFEngine = object\n")
    for name in ("ezdxf.py", "svgwrite.py"):
        open(os.path.join(STUB_DIR, name), "w").close()
    return True


def load():
    """Import the installed reference; returns dict of its hot-path modules or None if it is not installed."""
    if not os.path.exists(os.path.join(REF_DIR, "nearfield_farfield.py")):
        return None
    for p in (STUB_DIR, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import nearfield_farfield
        import nearfield
        import grating
        import lens_center
        import design_collimator
    return dict(nearfield_farfield=nearfield_farfield, nearfield=nearfield, grating=grating,
                lens_center=lens_center, design_collimator=design_collimator)


if __name__ == "__main__":
    print("installed" if install() else "reference tree not found")
