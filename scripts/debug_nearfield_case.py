"""Debug helper: run the 'larger lens' near-field case on the GPU and dump kernel vs oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
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
err = np.abs(got[0] - ref[0]) / np.abs(ref[0]).max()
bad = np.argwhere(err > 1e-9)
print("bad points:", len(bad), "of", err.size, "max", err.max())
r = np.hypot(*np.meshgrid(got[4], got[5], indexing="ij"))
for i, j in bad[:40]:
    print(i, j, "x=%.4e y=%.4e r=%.5e" % (got[4][i], got[5][j], r[i, j]), "got", got[0][i, j], "ref", ref[0][i, j])
print("r_min", periph["r_min_list"][:3], "r_max last", periph["r_max_list"][-1])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "nf_debug.npz"), bad=bad, err=err)
