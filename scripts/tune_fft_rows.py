"""Sweep the row-pass tuning knobs (mlb_fft_tune) on cfg3 and print the fused fold+FFT-rows time."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from metalens_b200 import _lib
from metalens_b200.farfield import FarfieldPlan
lib = _lib.load()
M, s = 4096, 4
wl, ng = 532e-9, 1.4607
d = wl / 2.2
g = torch.Generator(device="cuda").manual_seed(0)
fields = [[torch.randn(M, M, dtype=torch.complex64, device="cuda", generator=g) for _ in range(4)] for _ in range(3)]
plans = [FarfieldPlan((M, M), d, d, wl, ng, stride=s, method="fft") for _ in range(3)]
ref = None
def t_kernel(k):
    steps = [p.steps(f) for p, f in zip(plans, fields)]
    for i in range(6): steps[i % 3][k][1]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): steps[i % 3][k][1]()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30 * 1e3
for plain, pts, thr, vec in ((3, 1024, 256, 2), (2, 1024, 256, 2)):
    _lib.check(lib.mlb_fft_tune(plain, pts, thr, vec), "tune")
    try:
        us = t_kernel(0)
        plans[0].run(fields[0]); P = plans[0].P.clone()
        if ref is None: ref = P
        ok = bool(torch.allclose(torch.nan_to_num(P), torch.nan_to_num(ref), rtol=1e-4, atol=0))
        print("plain=%d points/cta=%d threads=%d vec=%d : rows %.1f us  (%.0f GB/s)  cols %.1f us  epilogue %.1f us  same=%s" % (plain, pts, thr, vec, us, 32 * (M * M + 1024 * 1024) / us / 1e3, t_kernel(1), t_kernel(2), ok), flush=True)
    except Exception as e:
        print("plain=%d pts=%d thr=%d vec=%d failed: %s" % (plain, pts, thr, vec, e), flush=True)
print("cols %.1f us" % t_kernel(1))
