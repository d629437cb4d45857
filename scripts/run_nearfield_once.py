"""Run the aperture-assembly kernel a few times on the bench lens (for ncu captures / quick timing).
usage: run_nearfield_once.py [M] [reps]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth_lens
from metalens_b200 import grating, lens_center
from metalens_b200.design import make_design
from metalens_b200.nearfield import NearfieldPlan
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wl = 580e-9; R = M * (wl / 2.2) / 2; f = R / math.tan(math.radians(44.0))
spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1)], source_distance=f, radius=R * 0.999)
collections, hgs = synth_lens.make_library(grating, lens_center, spec)
periph, center, _ = make_design(collections, f, spec["radius"], hgs)
plan = NearfieldPlan(wl, periph, center, hgs)
x = np.linspace(-R, R, M)
out = torch.zeros((4, M, M), dtype=torch.complex64, device="cuda")
for _ in range(2): plan.run(0.0, 0.0, -f, "x", x, x, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): plan.run(0.0, 0.0, -f, "x", x, x, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("nearfield %d^2: %.3f ms  %.2e samples/s  rings %d cells %d" % (M, ms, M * M / ms * 1e3, len(periph["r_min_list"]), len(center)))
