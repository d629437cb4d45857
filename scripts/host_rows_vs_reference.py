"""Dev container only (needs /root/reference): wall time of the host-side rows that surround the hot path --
build_interpolators() of GratingCollection / HexGridSet and make_design() (N2) -- in the unmodified
reference and in metalens_b200, on the synthetic lens library; results are checked equal.  CPU only.
usage: host_rows_vs_reference.py [lens radius in um, default 60]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import reference_loader
import synth_lens
ref = reference_loader.load()
from metalens_b200 import grating, lens_center
from metalens_b200.design import make_design

R = float(sys.argv[1]) * 1e-6 if len(sys.argv) > 1 else 60e-6
spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1)], source_distance=R / np.tan(np.radians(44.0)), radius=R)


def timed_library(gmod, lmod):
    t = {}
    cols = []
    t0 = time.perf_counter()
    for lo, hi, lot, salt in spec["bands"]:
        cols.append([(np.radians(lo), np.radians(hi)), synth_lens.make_collection(gmod, np.radians(lo), np.radians(hi), lot, salt=salt)])
    hgs = synth_lens.make_hexgridset(lmod, gmod)
    t["table synthesis (same code both sides)"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _, gc in cols:
        gc.build_interpolators()
    t["GratingCollection.build_interpolators x%d" % len(cols)] = time.perf_counter() - t0
    t0 = time.perf_counter()
    hgs.build_interpolators()
    t["HexGridSet.build_interpolators"] = time.perf_counter() - t0
    return cols, hgs, t


rc, rh, rt = timed_library(ref["grating"], ref["lens_center"])
oc, oh, ot = timed_library(grating, lens_center)
t0 = time.perf_counter()
rp, rcen, _ = ref["design_collimator"].make_design(rc, spec["source_distance"], spec["radius"], rh)
rt["make_design (R = %.0f um)" % (R * 1e6)] = time.perf_counter() - t0
t0 = time.perf_counter()
op, ocen, _ = make_design(oc, spec["source_distance"], spec["radius"], oh)
ot["make_design (R = %.0f um)" % (R * 1e6)] = time.perf_counter() - t0
assert np.array_equal(np.asarray(rcen), np.asarray(ocen)), "centre cells differ"
for k in ("r_min_list", "r_max_list", "r_center_list", "grating_period_list", "num_around_circle_list"):
    assert np.array_equal(np.asarray(rp[k]), np.asarray(op[k])), k
# the tables themselves (the callable of metalens_b200 evaluates on the GPU; here only grids and values are compared)
for key in sorted(rc[0][1].interpolators, key=repr):
    a, b = rc[0][1].interpolators[key], oc[0][1].interpolators[key]
    assert all(np.array_equal(g1, g2) for g1, g2 in zip(a.grid, b.grid)) and np.array_equal(a.values, b.values), key
assert tuple(rc[0][1].interpolator_bounds) == tuple(oc[0][1].interpolator_bounds)
for key in sorted(rh.interpolators, key=repr):
    assert np.array_equal(rh.interpolators[key].values, oh.interpolators[key].values), key
print("lens: %d rings, %d hex cells" % (len(op["r_min_list"]), len(ocen)))
for k in rt:
    print("%-48s reference %8.3f s   metalens_b200 %8.3f s   x%.1f" % (k, rt[k], ot[k], rt[k] / max(ot[k], 1e-9)))
