"""Which samples of the 225 x 225 odd-grid lens differ between build_nearfield and the oracle, by class."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth_lens
from oracle import nearfield_oracle as no
from metalens_b200 import grating, lens_center
from metalens_b200.design import make_design
from metalens_b200.nearfield import build_nearfield

spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1)], source_distance=30e-6, radius=29e-6)
collections, hgs = synth_lens.make_library(grating, lens_center, spec)
periph, center, r_switch = make_design(collections, spec["source_distance"], spec["radius"], hgs)
args = (1.1e-6, -0.6e-6, -30e-6, "z", 580e-9, periph, center, hgs)
got = build_nearfield(*args)
ref = no.build_nearfield(*args)
X, Y = np.meshgrid(got[4], got[5], indexing="ij")
r = np.hypot(X, Y)
d2 = (X.ravel()[:, None] - center[None, :, 0]) ** 2 + (Y.ravel()[:, None] - center[None, :, 1]) ** 2
in_center = r.ravel() <= periph["r_min_list"][0]
tied = (((d2 <= d2.min(axis=1, keepdims=True)).sum(axis=1) > 1) & in_center).reshape(X.shape)
ring = np.searchsorted(np.hstack((periph["r_min_list"], periph["r_max_list"][-1])), r) - 1
ring[ring == len(periph["r_min_list"])] = -1
apg = 2 * np.pi / periph["num_around_circle_list"][np.maximum(ring, 0)]
turns = np.arctan2(Y, X) / apg
wedge = (np.abs(np.abs(turns - np.round(turns)) - 0.5) < 1e-9) & (ring >= 0)
bad = np.zeros(X.shape, bool)
for k in range(4):
    bad |= np.abs(got[k] - ref[k]) > 1e-9 * np.abs(ref[k]).max()
print("bad", bad.sum(), "tied", tied.sum(), "wedge", wedge.sum(), "bad&tied", (bad & tied).sum(), "bad&wedge", (bad & wedge).sum(),
      "bad other", (bad & ~tied & ~wedge).sum())
for i, j in list(zip(*np.nonzero(bad)))[:12]:
    print(i, j, "x %.17g y %.17g ring %d turns %.17g frac-0.5 %.3e tied %s" % (X[i, j], Y[i, j], ring[i, j], turns[i, j],
          abs(turns[i, j] - np.round(turns[i, j])) - 0.5, tied[i, j]))
