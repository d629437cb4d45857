"""cfg3 step (3 x 4096^2 -> 1024^2, two-stream pipeline) under the tuning options of the fused column+power pass."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import apertures
from metalens_b200 import _lib
from metalens_b200.farfield import FarfieldPlan
from metalens_b200.sharding import ShardedFarfield
lib = _lib.load()
M, s = 4096, 4
items = [(450e-9, 1.466), (532e-9, 1.4607), (635e-9, 1.457)]
fields = {}
for i, (wl, ng) in enumerate(items):
    a = apertures.focusing_lens(M, 1000 + i, wl, ng)
    fields[i] = [torch.from_numpy(v).cuda() for v in a[:4]]
d = float(apertures.grid(M, 532e-9)[0][1] - apertures.grid(M, 532e-9)[0][0])

def build():
    def make_plan(item, r0, r1):
        wl, ng = items[item]
        x = apertures.grid(M, wl)[0]
        return FarfieldPlan((M, M), float(x[1] - x[0]), float(x[1] - x[0]), wl, ng, stride=s)
    return ShardedFarfield(3, M // s, make_plan, rank=0, world=1)

def timed(fn, n=30, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

ref = None
for narrow in (0, 1):
    for per_sm in (0, 1):
        lib.mlb_set_option(b"cols_power_narrow", narrow)
        lib.mlb_set_option(b"rows_ctas_per_sm", per_sm)
        sh = build()
        def step():
            sh.run(lambda i: fields[i], overlap=True)
        us = timed(lambda: (step(), None)[1])
        sh.finish(); torch.cuda.synchronize()
        P = sh.run(lambda i: fields[i])[0].clone()
        if ref is None: ref = P
        same = bool(((P == ref) | (torch.isnan(P) & torch.isnan(ref))).all())
        st = sh.plans[0].steps(fields[0])
        k_us = {name: timed(fn, 20, 3) for name, fn, _b, _f in st}
        print(json.dumps(dict(narrow=narrow, rows_ctas_per_sm=per_sm, step_us=us, same_as_default=same, kernels_us=k_us)), flush=True)
lib.mlb_set_option(b"cols_power_narrow", 0); lib.mlb_set_option(b"rows_ctas_per_sm", 0)
