"""Kernel end times of the two-stream pipelined step (cfg3 shape) for different stream priorities and
column-tile widths: prints, per configuration, when each aperture pass (main stream) and each tail
(side stream) finished relative to the start of the measured window.  Run under gpurun."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from metalens_b200 import _lib  # noqa: E402
from metalens_b200.farfield import FarfieldPlan  # noqa: E402

lib = _lib.load()
wl, ng = 532e-9, 1.4607
d = wl / 2.2
M, s, n_items = 4096, 4, 3
g = torch.Generator(device="cuda").manual_seed(0)
fields = [[torch.randn(M, M, dtype=torch.complex64, device="cuda", generator=g) for _ in range(4)] for _ in range(n_items)]
plans = [FarfieldPlan((M, M), d, d, wl, ng, stride=s) for _ in range(n_items)]


def run(prio_rows, prio_tail, wide, steps=4, per_sm=0, ring=64):
    lib.mlb_set_option(b"rows_ring_kb", ring)
    lib.mlb_set_option(b"cols_power_wide", wide)
    lib.mlb_set_option(b"rows_ctas_per_sm", per_sm)
    A = torch.cuda.Stream(priority=prio_rows)
    B = torch.cuda.Stream(priority=prio_tail)
    tail_done = [None] * n_items
    marks = []

    def one_step(record):
        for k, p in enumerate(plans):
            first, second = p.run_split(fields[k])
            with torch.cuda.stream(A):
                if tail_done[k] is not None:
                    A.wait_event(tail_done[k])
                first()
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(A)
            with torch.cuda.stream(B):
                B.wait_event(ev)
                second()
                dn = torch.cuda.Event(enable_timing=True)
                dn.record(B)
                tail_done[k] = dn
            if record:
                marks.append((k, ev, dn))
    for _ in range(3):
        one_step(False)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(A):
        e0.record(A)
    for _ in range(steps):
        one_step(True)
    torch.cuda.synchronize()
    rows_end = [round(e0.elapsed_time(ev) * 1e3, 1) for _, ev, _ in marks]
    tail_end = [round(e0.elapsed_time(dn) * 1e3, 1) for _, _, dn in marks]
    total = max(rows_end[-1], tail_end[-1])
    print(json.dumps(dict(prio_rows=prio_rows, prio_tail=prio_tail, wide=wide, rows_per_sm=per_sm, ring=ring,
                          us_per_item=round(total / (steps * n_items), 1), rows_end=rows_end, tail_end=tail_end)), flush=True)


print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "n/a", flush=True)
for ring in (128, 64):
    for per_sm in (1, 0):
        for wide in (0, 1):
            for pr, pt in ((-1, 0), (0, 0)):
                if ring == 64 and per_sm == 0 and (pr, pt) == (-1, 0):
                    continue
                run(pr, pt, wide, per_sm=per_sm, ring=ring, steps=3)
lib.mlb_set_option(b"cols_power_wide", -1)
lib.mlb_set_option(b"rows_ctas_per_sm", 0)
lib.mlb_set_option(b"rows_ring_kb", 64)
