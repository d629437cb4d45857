"""Golden amplitude tables: GratingCollection.build_interpolators / HexGridSet.build_interpolators of the UNMODIFIED
reference (grating.py:1186-1232, lens_center.py:188-226) on the synthetic SMALL_LENS library.  Dev container only.

    python tests/golden/make_tables_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import reference_loader  # noqa: E402
import synth_lens  # noqa: E402

ref = reference_loader.load()
collections, hgs = synth_lens.make_library(ref["grating"], ref["lens_center"], synth_lens.SMALL_LENS)
out = {}
for name, owner in [("gc0", collections[0][1]), ("gc1", collections[1][1]), ("hgs", hgs)]:
    keys = sorted(owner.interpolators, key=repr)
    out[name + "_keys"] = np.array([repr(k) for k in keys])
    out[name + "_values"] = np.stack([owner.interpolators[k].values for k in keys])
    g = owner.interpolators[keys[0]].grid
    for a in range(3):
        out["%s_grid%d" % (name, a)] = np.asarray(g[a], dtype=np.float64)
    out[name + "_bounds"] = np.asarray(owner.interpolator_bounds, dtype=np.float64)
path = os.path.join(HERE, "tables_small_lens.npz")
np.savez_compressed(path, **out)
print({k: v.shape for k, v in out.items() if k.endswith("values")}, "%.2f MB" % (os.path.getsize(path) / 1e6))
