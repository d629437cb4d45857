"""All-bins (stride 1: the reference's own grid, nearfield_farfield.py:35-39) NF->FF, per-kernel timings for a
list of aperture sizes; random fields (timing only).  usage: allbins_kernels.py [sizes...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from metalens_b200 import _lib
from metalens_b200.farfield import FarfieldPlan
lib = _lib.load()
engine = int(os.environ.get("ROWS_ENGINE", "0"))
lib.mlb_set_option(b"rows_engine", engine)
cengine = int(os.environ.get("COLS_ENGINE", "0"))
lib.mlb_set_option(b"cols_engine", cengine)
occ = int(os.environ.get("R16_OCC", "0"))
lib.mlb_set_option(b"r16_occupancy", occ)
strip = int(os.environ.get("COLS_STRIP_MB", "48"))
lib.mlb_set_option(b"cols_strip_mb", strip)
lib.mlb_set_option(b"mixed_registers", int(os.environ.get("MIXED_REG", "1")))
lib.mlb_set_option(b"mixed_occupancy", int(os.environ.get("MIXED_OCC", "0")))
print("mixed_registers", lib.mlb_get_option(b"mixed_registers"), "mixed_occupancy", lib.mlb_get_option(b"mixed_occupancy"))
fuse_mode = os.environ.get("FUSE", "default")          # default | always | never
print("rows_engine", engine, "cols_engine", cengine, "r16_occupancy", occ, "cols_strip_mb", strip, "fuse", fuse_mode)
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 3375, 3600, 4096, 8192]
wl, ng = 580e-9, 1.459
d = wl / 2.2
HBM = 6541.8


def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for M in sizes:
    g = torch.Generator(device="cuda").manual_seed(0)
    # two field sets so that consecutive launches never find their input in L2 (for M >= 2048)
    sets = [[torch.randn(M, M, dtype=torch.complex64, device="cuda", generator=g) for _ in range(4)] for _ in range(2)]
    for fuse in (True, False):
        if (fuse_mode == "never" and fuse) or (fuse_mode == "always" and not fuse):
            continue
        plan = FarfieldPlan((M, M), d, d, wl, ng, stride=1, fuse_power=("always" if fuse_mode == "always" else fuse))
        if fuse and not plan.fused:
            continue
        rr = [0]
        def step():
            rr[0] ^= 1
            plan.run(sets[rr[0]])
        t = timed(step)
        ideal = (32 * M * M * 4 + 4 * M * M) / HBM / 1e6          # rows r+w, cols r, P w  (ms)
        print("M=%d all bins (%s%s): %.3f ms = %.2e points/s; HBM floor (2 passes) %.3f ms -> %.0f %%" % (
            M, plan.method, ", fused" if plan.fused else "", t, M * M / t * 1e3, ideal, 100 * ideal / t), flush=True)
        st = [plan.steps(s) for s in sets]
        for k in range(len(st[0])):
            def one(k=k):
                rr[0] ^= 1
                st[rr[0]][k][1]()
            tk = timed(one)
            print("    %-16s %.3f ms  %.0f GB/s (algorithmic)" % (st[0][k][0], tk, st[0][k][2] / tk / 1e6), flush=True)
        del plan, st
    del sets
    torch.cuda.empty_cache()
