"""Time the aperture-assembly kernel variants (mlb_nearfield_tune) on the bench lens."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth_lens
from metalens_b200 import _lib, grating, lens_center
from metalens_b200.design import make_design
from metalens_b200.nearfield import NearfieldPlan
lib = _lib.load()
M = 4096; wl = 580e-9; R = M * (wl / 2.2) / 2; f = R / math.tan(math.radians(44.0))
spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1)], source_distance=f, radius=R * 0.999)
collections, hgs = synth_lens.make_library(grating, lens_center, spec)
periph, center, _ = make_design(collections, f, spec["radius"], hgs)
plan = NearfieldPlan(wl, periph, center, hgs)
x = np.linspace(-R, R, M)
ref = None
for lg in (5, 4, 3, 2):          # warp tile 32x1, 16x2, 8x4, 4x8 (y x x)
    _lib.check(lib.mlb_nearfield_tune(100 + lg), "tile")
    _lib.check(lib.mlb_nearfield_tune(6), "tune")
    out = torch.zeros((4, M, M), dtype=torch.complex64, device="cuda")
    for _ in range(2): plan.run(0.0, 0.0, -f, "x", x, x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _o, pw = plan.run(0.0, 0.0, -f, "x", x, x, out=out)
    e1.record(); torch.cuda.synchronize()
    if lg == 5: ref5, pw5 = out.clone(), pw.item()
    print("warp tile %2d x %d: %.3f ms  max|diff| vs 32x1 %.2e  power rel diff %.1e" % (1 << lg, 32 >> lg, e0.elapsed_time(e1) / 5,
          (out - ref5).abs().max().item() / ref5.abs().max().item(), abs(pw.item() - pw5) / abs(pw5)), flush=True)
_lib.check(lib.mlb_nearfield_tune(103), "tile")
for variant, dtype in ((1, torch.complex64), (5, torch.complex64), (6, torch.complex64), (8, torch.complex64), (1, torch.complex128)):
    _lib.check(lib.mlb_nearfield_tune(variant), "tune")
    out = torch.zeros((4, M, M), dtype=dtype, device="cuda")
    for _ in range(2): plan.run(0.0, 0.0, -f, "x", x, x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): plan.run(0.0, 0.0, -f, "x", x, x, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if dtype == torch.complex128:
        err = ((ref.to(torch.complex128) - out).abs().max() / out.abs().max()).item()
        print("complex128 kernel: %.3f ms ; complex64 (variant 1) vs complex128 max rel err %.2e" % (ms, err))
    else:
        if ref is None: ref = out.clone()
        print("variant minBlocks=%d complex64: %.3f ms (%.2e samples/s)  identical to variant 1: %s" % (variant, ms, M * M / ms * 1e3, bool(torch.equal(ref, out))))
