"""A/B: pipelined cfg3 / cfg2 step (ShardedFarfield, two streams) with the row pass drawing its rows from a work
counter (rows_dynamic 1) or by a fixed stride (0); CUDA events, device-resident inputs.  usage: ab_rows_dynamic.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import apertures
from metalens_b200 import _lib
from metalens_b200.farfield import FarfieldPlan
from metalens_b200.sharding import ShardedFarfield
lib = _lib.load()
for tag, M, s, items in (("cfg3", 4096, 4, [(450e-9, 1.466), (532e-9, 1.4607), (635e-9, 1.457)]),
                         ("cfg2", 2048, 4, [(532e-9, 1.4607), (532e-9, 1.4607)])):
    K = M // s
    g = torch.Generator(device="cuda").manual_seed(1)
    fields = [[torch.randn(M, M, dtype=torch.complex64, device="cuda", generator=g) for _ in range(4)] for _ in items]
    def make_plan(item, r0, r1):
        wl, ng = items[item]
        d = wl / 2.2
        return FarfieldPlan((M, M), d, d, wl, ng, stride=s)
    ref = None
    for dyn, prio in ((0, 0), (1, 0), (0, -1), (1, -1), (0, 0), (1, -1)):
        sh = ShardedFarfield(len(items), K, make_plan, rank=0, world=1, tail_priority=prio)
        lib.mlb_set_option(b"rows_dynamic", dyn)
        for _ in range(5): sh.run(lambda i: fields[i], overlap=True)
        sh.finish(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 50
        e0.record()
        for _ in range(steps): res = sh.run(lambda i: fields[i], overlap=True)
        sh.finish(); e1.record(); torch.cuda.synchronize()
        P = res[0].clone()
        if ref is None: ref = P
        same = bool(((P == ref) | (torch.isnan(P) & torch.isnan(ref))).all())
        us = e0.elapsed_time(e1) / steps * 1e3
        # rows kernel alone
        plan = sh.plans[0]
        st = plan.steps(fields[0])
        for _ in range(3): st[0][1]()
        torch.cuda.synchronize(); e0.record()
        for k in range(20): sh.plans[k % len(items)].steps(fields[k % len(items)])[0][1]()
        e1.record(); torch.cuda.synchronize()
        print("%s rows_dynamic=%d tail_priority=%d: %.1f us/step (%.3e pts/s), rows kernel alone %.1f us, identical to first: %s"
              % (tag, dyn, prio, us, len(items) * K * K / us * 1e6, e0.elapsed_time(e1) / 20 * 1e3, same), flush=True)
        del sh
    del fields
    torch.cuda.empty_cache()
lib.mlb_set_option(b"rows_dynamic", 1)
