"""A/B of the big-radix mixed engine's register kernels: compile-time plans (mixed_compiled = 1) against the generic
run-time-plan kernels (0), per pass, on all-bins NF->FF at good_fft_number() sizes; random fields (timing only).
usage: mixed_ab.py [out.json] [sizes...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from metalens_b200 import _lib
from metalens_b200.farfield import FarfieldPlan

lib = _lib.load()
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "mixed_ab.json")
sizes = [int(a) for a in sys.argv[2:]] or [3375, 2700, 3600, 2160]
wl, ng = 580e-9, 1.459
d = wl / 2.2


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = []
for M in sizes:
    g = torch.Generator(device="cuda").manual_seed(0)
    sets = [[torch.randn((M, M + (M & 1)), generator=g, device="cuda", dtype=torch.float32).to(torch.complex64)[:, :M]
             for _ in range(4)] for _ in range(2)]              # two sets: consecutive launches never find their input in L2
    plan = FarfieldPlan((M, M), d, d, wl, ng, stride=1, method="fft")
    for compiled, occ in ((0, 0), (1, 0), (1, 3)):
        lib.mlb_set_option(b"mixed_compiled", compiled)
        lib.mlb_set_option(b"mixed_occupancy", occ)
        row = {"M": M, "mixed_compiled": compiled, "mixed_occupancy": occ}
        steps = [plan.steps(f) for f in sets]
        k = [0]

        def whole():
            k[0] ^= 1
            for _n, fn, _b, _f in steps[k[0]]:
                fn()
        row["total_ms"] = timed(whole)
        for i, (name, _fn, nbytes, _fl) in enumerate(steps[0]):
            def one(i=i):
                k[0] ^= 1
                steps[k[0]][i][1]()
            ms = timed(one)
            row[name + "_ms"] = ms
            row[name + "_gbs"] = nbytes / ms * 1e-6
        P, total = plan.run(sets[0])
        row["total_P"] = float(total.item())
        res.append(row)
        print(json.dumps(row), flush=True)
    del plan, sets
    torch.cuda.empty_cache()
lib.mlb_set_option(b"mixed_compiled", 1)
lib.mlb_set_option(b"mixed_occupancy", 0)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
with open(out_path, "w") as f:
    json.dump(res, f, indent=1)
