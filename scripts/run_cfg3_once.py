"""Run a few cfg3-shaped NF->FF items (sequential, fused) -- a short target for ncu captures.
usage: run_cfg3_once.py [M] [stride] [wide(-1/0/1)]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from metalens_b200 import _lib
from metalens_b200.farfield import FarfieldPlan
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = int(sys.argv[2]) if len(sys.argv) > 2 else 4
wide = int(sys.argv[3]) if len(sys.argv) > 3 else -1
lib = _lib.load()
lib.mlb_set_option(b"cols_power_wide", wide)
wl, ng = 532e-9, 1.4607
d = wl / 2.2
g = torch.Generator(device="cuda").manual_seed(0)
fields = [[torch.randn(M, M, dtype=torch.complex64, device="cuda", generator=g) for _ in range(4)] for _ in range(2)]
plans = [FarfieldPlan((M, M), d, d, wl, ng, stride=s) for _ in range(2)]
for it in range(3):
    for p, f in zip(plans, fields):
        p.run(f)
torch.cuda.synchronize()
print("done", float(plans[0].total))
