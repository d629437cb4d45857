"""Measure the NF->FF step variants on one B200 (run under gpurun; prints one JSON line per measurement).

  seq-unfused   rows, cols, epilogue, one stream            (round-1 baseline)
  seq-fused     rows, fused cols+power, one stream
  pipe          two-stream pipeline of ShardedFarfield.run(overlap=True): the column/power tail of item k
                under the HBM-bound aperture pass of item k+1; swept over the row-pass options
  graph         the pipelined step captured in one CUDA graph
plus the kernels timed alone.  Synthetic random apertures (timing only)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from metalens_b200 import _lib  # noqa: E402
from metalens_b200.farfield import FarfieldPlan  # noqa: E402
from metalens_b200.sharding import ShardedFarfield  # noqa: E402

lib = _lib.load()
wl, ng = 532e-9, 1.4607
d = wl / 2.2


def opt(**kw):
    for k, v in kw.items():
        _lib.check(lib.mlb_set_option(k.encode(), int(v)), k)


def out(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, fin, steps=20, warm=4):
    for _ in range(warm):
        fn()
    fin()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    t_issue = time.perf_counter() - t0
    fin()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3, t_issue / steps * 1e6     # us per step (device), host issue us


def workload(M, s, n_items, tag):
    K = M // s
    g = torch.Generator(device="cuda").manual_seed(0)
    fields = [[torch.randn(M, M, dtype=torch.complex64, device="cuda", generator=g) for _ in range(4)]
              for _ in range(n_items)]

    def sharded(fuse):
        return ShardedFarfield(n_items, K, lambda i, r0, r1: FarfieldPlan((M, M), d, d, wl, ng, stride=s,
                                                                            fuse_power=fuse), rank=0, world=1)
    nop = lambda: None
    opt(rows_ctas_per_sm=0, rows_l2_evict_first=1, cols_power_wide=0)
    sh_u, sh_f = sharded(False), sharded(True)
    us, host = timed(lambda: sh_u.run(lambda i: fields[i]), nop)
    out(tag=tag, variant="seq-unfused", us_per_step=us, host_us=host, us_per_item=us / n_items)
    for wide in (0, 1):
        opt(cols_power_wide=wide)
        us, host = timed(lambda: sh_f.run(lambda i: fields[i]), nop)
        out(tag=tag, variant="seq-fused", wide=wide, us_per_step=us, host_us=host, us_per_item=us / n_items)
    # kernels alone (rotating over the items so the aperture never sits in L2)
    for wide in (0, 1):
        for evict in (1, 0):
            for per_sm in (0, 1):
                opt(cols_power_wide=wide, rows_l2_evict_first=evict, rows_ctas_per_sm=per_sm)
                st = [p.steps(fields[t.item]) for t, p in zip(sh_f.tiles, sh_f.plans)]
                for p, t in zip(sh_f.plans, sh_f.tiles):
                    p.run(fields[t.item])
                rr = [0]
                res = {}
                for k in range(len(st[0])):
                    def one(k=k):
                        rr[0] = (rr[0] + 1) % len(st)
                        st[rr[0]][k][1]()
                    res[st[0][k][0]] = timed(one, nop, steps=30, warm=6)[0]
                out(tag=tag, variant="kernels", wide=wide, evict_first=evict, rows_per_sm=per_sm, us=res)
    st = [p.steps(fields[t.item]) for t, p in zip(sh_u.tiles, sh_u.plans)]
    res = {}
    rr = [0]
    for k in range(len(st[0])):
        def one(k=k):
            rr[0] = (rr[0] + 1) % len(st)
            st[rr[0]][k][1]()
        res[st[0][k][0]] = timed(one, nop, steps=30, warm=6)[0]
    out(tag=tag, variant="kernels-unfused", us=res)
    # two-stream pipeline
    best = None
    for fuse, sh in ((True, sh_f), (False, sh_u)):
        for wide in ((0, 1) if fuse else (0,)):
            for evict in (1, 0):
                for per_sm in (0, 1):
                    opt(cols_power_wide=wide, rows_l2_evict_first=evict, rows_ctas_per_sm=per_sm)
                    us, host = timed(lambda: sh.run(lambda i: fields[i], overlap=True), sh.finish)
                    out(tag=tag, variant="pipe", fused=fuse, wide=wide, evict_first=evict, rows_per_sm=per_sm,
                        us_per_step=us, host_us=host, us_per_item=us / n_items)
                    if fuse and (best is None or us < best[0]):
                        best = (us, wide, evict, per_sm)
    _, wide, evict, per_sm = best
    opt(cols_power_wide=wide, rows_l2_evict_first=evict, rows_ctas_per_sm=per_sm)
    try:
        sh_f.capture(lambda i: fields[i])
        us, host = timed(sh_f.replay, nop)
        out(tag=tag, variant="graph", wide=wide, evict_first=evict, rows_per_sm=per_sm, us_per_step=us, host_us=host,
            us_per_item=us / n_items)
    except Exception as e:                                  # noqa: BLE001
        out(tag=tag, variant="graph", error=repr(e))
    opt(rows_ctas_per_sm=0, rows_l2_evict_first=1, cols_power_wide=-1)
    del fields, sh_u, sh_f
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg2", "cfg4s"]
    if "cfg3" in which:
        workload(4096, 4, 3, "cfg3 4096^2->1024^2 x3")
    if "cfg2" in which:
        workload(2048, 4, 2, "cfg2 2048^2->512^2 x2")
    if "cfg4s" in which:
        workload(8192, 4, 2, "cfg4 8192^2->2048^2 x2")
