#!/usr/bin/env python
"""bench.py -- far-field points/s of the NF->FF hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2]

One "step" = one pass of the hot path over one batch: every batch item (wavelength /
polarisation) of the workload goes aperture fields -> far-field power map P.

  cfg3 (default; the configuration the north-star target is quoted on):
       4096^2 aperture -> 1024^2 far field (every 4th fftshifted FFT bin), 3 wavelengths
  cfg2 2048^2 aperture -> 512^2 far field, 532 nm, TE+TM (2 items)

value  = far-field points/s with the aperture fields already resident in HBM
e2e    = same metric through FarfieldPlan.run_host(): pinned host fields -> H2D ->
         kernels -> D2H of P, every step
N > 1  : one process per GPU (torchrun), weak scaling -- every rank owns a full batch of its
         own apertures (far-field tiles of different sources); one all-gather of the P tiles per
         step is inside the timed region (asynchronous, double-buffered: it overlaps the kernels of
         the next step).  The gather pulls the peers' tiles over NVLink with the copy engines
         (torch symmetric memory; no SM-resident collective kernel next to the persistent row
         pass) and falls back to NCCL all_gather_into_tensor where that cannot be set up.
Extra keys on the same line: roofline (dominant kernel), kernels (CUDA-event time of every kernel of a
step), paths_points_per_s (other formulations on the same workload), other_workloads (cfg2),
nearfield_assembly (hot path B), fom_sweep (cfg5 shape), cpu_baseline.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import apertures  # noqa: E402

WORKLOADS = {
    "cfg3": dict(M=4096, stride=4, items=[(450e-9, 1.466, False), (532e-9, 1.4607, False), (635e-9, 1.457, False)],
                 name="cfg3: 4096x4096 aperture -> 1024x1024 far field (every 4th FFT bin), 450/532/635 nm"),
    "cfg2": dict(M=2048, stride=4, items=[(532e-9, 1.4607, False), (532e-9, 1.4607, True)],
                 name="cfg2: 2048x2048 aperture -> 512x512 far field (every 4th FFT bin), 532 nm, TE+TM"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference(workload, sample_M, steps=1, warmup=0, parallel=True):
    """Time the CPU reference path (4x fft2(fftshift) + radiated-power helper, float64 numpy: the oracle's
    restatement of nearfield_farfield.py) on a bounded sample: the same workload with the aperture reduced to
    sample_M^2 samples (stride kept), all batch items.  All host cores are used: the items run concurrently and
    each spreads its four FFTs and the reference's own uy-chunk loop (:45-66) over its share of the cores
    (threads; numpy releases the GIL).  The timed region is the transform only (input synthesis excluded)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import farfield_oracle as fo
    w = WORKLOADS[workload]
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = min(avail, 64) if parallel else 1          # 64: bounds the chunk temporaries on many-core hosts
    n_items = len(w["items"])
    per_item = max(1, -(-cores // n_items))
    K = sample_M // w["stride"]
    inputs = [apertures.focusing_lens(sample_M, 100 + i, wl, ng, rotate=rot) + (wl, ng)
              for i, (wl, ng, rot) in enumerate(w["items"])]

    def one(a):
        return float(fo.farfield_reference_path_threads(*a, workers=per_item)[1])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if cores > 1:
            with ThreadPoolExecutor(n_items) as pool:
                list(pool.map(one, inputs))
        else:
            for a in inputs:
                fo.farfield_reference_path(*a)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    pts = n_items * K * K
    used = min(cores, n_items * per_item)
    return dict(value=pts / t, unit="far-field points/s", cores=used, kind="port",
                sample="%d items, aperture %dx%d -> %dx%d requested bins (the numpy path computes all %d^2 bins: "
                       "%.3g bins/s); oracle/farfield_oracle.py restating nearfield_farfield.py, float64, "
                       "%d thread(s) (items concurrent, FFTs and the uy-chunk loop threaded), %.2f s per pass"
                       % (n_items, sample_M, sample_M, K, K, sample_M, n_items * sample_M ** 2 / t, used, t)), t


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is
    pure Python and cannot travel to the GPU box) on all host cores (cpu_reference)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    sample_M = min(w["M"], 2048)
    base, t = cpu_reference(args.workload, sample_M, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "far-field points/sec (NF->FF)", "value": base["value"],
        "unit": "far-field points/s", "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)),
        "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "sample": base["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "far-field points/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------- hot path B
def bench_nearfield(M, torch, cpu_cols=384):
    """Aperture-field assembly (build_nearfield) on a synthetic round lens filling an M x M grid
    (SURVEY 8d cfg4 shape): samples/s and achieved write bandwidth of the fused kernel, next to the
    numpy oracle timed on a strip of the same grid."""
    import synth_lens
    from metalens_b200 import grating, lens_center
    from metalens_b200.design import make_design
    from metalens_b200.nearfield import NearfieldPlan
    wl = 580e-9
    R = M * (wl / 2.2) / 2
    f = R / math.tan(math.radians(44.0))
    spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1)], source_distance=f, radius=R * 0.999)
    t0 = time.perf_counter()
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph, center, _ = make_design(collections, f, spec["radius"], hgs)
    t_design = time.perf_counter() - t0
    plan = NearfieldPlan(wl, periph, center, hgs)
    x = np.linspace(-R, R, M)
    out = torch.zeros((4, M, M), dtype=torch.complex64, device="cuda")
    for _ in range(2):
        plan.run(0.0, 0.0, -f, "x", x, x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        plan.run(0.0, 0.0, -f, "x", x, x, out=out)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    res = dict(aperture=[M, M], rings=int(len(periph["r_min_list"])), hex_cells=int(len(center)),
               ms=t * 1e3, samples_per_s=M * M / t, write_gbs=32.0 * M * M / t / 1e9,
               note="one fused kernel launch incl. host packing of x/y and the violation read-back; "
                    "algorithmic bytes = 32*M^2 written (4 complex64 fields)", design_seconds=t_design)
    # CPU oracle on a strip of the same grid (off-centre so centre and rings are both represented)
    from oracle import nearfield_oracle as no
    j0 = M // 2 + M // 8
    ys = x[j0:j0 + cpu_cols]
    t0 = time.perf_counter()
    no.build_nearfield(0.0, 0.0, -f, "x", wl, periph, center, hgs, x_pts=x, y_pts=ys)
    tc = time.perf_counter() - t0
    res["cpu_baseline"] = dict(value=M * cpu_cols / tc, unit="aperture samples/s", cores=1, kind="port",
                               sample="%d x %d strip of the same grid, oracle/nearfield_oracle.py (numpy + scipy "
                                      "cKDTree), 1 thread, %.1f s" % (M, cpu_cols, tc))
    return res


def bench_sweep(torch, steps=56, M=256):
    """BASELINE config 5 shape: a 5..60 degree sweep of 256 x 256 beam-deflector patches, one NF->FF
    + figure-of-merit evaluation per step; eager launches vs one captured CUDA graph replayed per step."""
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.fom import FarfieldFOM
    wl, ng = 532e-9, apertures.N_GLASS[532]
    plan = FarfieldPlan((M, M), wl / 2.2, wl / 2.2, wl, ng, stride=1)
    fom = FarfieldFOM(plan, 0.3, 0.0, 0.05)
    angles = np.linspace(5.0, 60.0, steps)
    patches = []
    for a in angles[:8]:
        Ex, Ey, Hx, Hy, x, y = apertures.tilted_te(M, wl, ng, angle_deg=float(a))
        patches.append(torch.stack([torch.from_numpy(v.astype(np.complex64)) for v in (Ex, Ey, Hx, Hy)]).cuda())

    def run(replay):
        vals = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            fom.static_fields[:, :, :M].copy_(patches[k % len(patches)])      # the step's new aperture
            r = fom.replay() if replay else fom.evaluate()
            vals.append(r.clone())
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps * 1e6, vals
    fom.evaluate()
    run(False)
    eager_us, v0 = run(False)
    fom.capture()
    run(True)
    graph_us, v1 = run(True)
    same = all(bool(torch.equal(a, b)) for a, b in zip(v0, v1))
    return dict(steps=steps, aperture=[M, M], far_field=[plan.Kx, plan.Ky], method=plan.method,
                eager_us_per_step=eager_us, graph_us_per_step=graph_us, identical_results=same,
                note="wall clock per step incl. the device copy of the step's aperture; 56 angles 5..60 deg")


# ----------------------------------------------------------------------------- GPU arm
def ours(args):
    import torch
    import torch.distributed as dist
    from metalens_b200 import _lib
    from metalens_b200.farfield import FarfieldPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first (rank 0, N=1 only), before the GPU work so that the host cores are idle
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        wl_M = WORKLOADS[args.workload]["M"]
        cpu, _ = cpu_reference(args.workload, min(wl_M, 2048) if not args.quick_cpu else 512, steps=1, warmup=0)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    w = WORKLOADS[args.workload]
    M, s = w["M"], w["stride"]
    K = M // s
    n_items = len(w["items"])
    peaks = measured_peaks()

    # far-field tiles over the ranks (metalens_b200/sharding.py): weak scaling -> world x n_items batch
    # items, whole items per rank, so each rank synthesises and owns only its own apertures
    from metalens_b200.sharding import ShardedFarfield
    n_global = n_items * world
    geom = {}

    def make_plan(item, r0, r1):
        wl, ng, rot = w["items"][item % n_items]
        d = apertures.grid(M, wl)[0]
        d = float(d[1] - d[0])
        rows = None if (r0, r1) == (0, K) else (r0, r1)
        return FarfieldPlan((M, M), d, d, wl, ng, stride=s, method=args.method if rows is None else "fold", rows=rows)

    sharded = ShardedFarfield(n_global, K, make_plan, rank=rank, world=world)
    plans = sharded.plans
    pinned, dev_fields = {}, {}
    for item in sharded.items_needed:
        wl, ng, rot = w["items"][item % n_items]
        Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 1000 + item, wl, ng, rotate=rot)
        pin = torch.empty((4, M, M), dtype=torch.complex64).pin_memory()
        for f, a in enumerate((Ex, Ey, Hx, Hy)):
            pin[f].copy_(torch.from_numpy(a))
        pinned[item] = pin
        dev_fields[item] = pin.cuda()
        del Ex, Ey, Hx, Hy

    def fields_of(item):
        return [dev_fields[item][f] for f in range(4)]

    def step_device():
        # local tiles + the one all-gather, which overlaps the next step's kernels (double-buffered)
        return sharded.run(fields_of, overlap=True)

    def host_runner(plan, pin):
        plan.run_host(pin)                       # pinned host -> H2D -> kernels -> D2H of P and total_P
        return plan.P, plan.total

    def step_host():
        return sharded.run(lambda item: pinned[item], runner=host_runner)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
            time.sleep(0.25)
        l0 = lib.mlb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        sharded.finish()                          # outstanding asynchronous all-gathers
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        launches = lib.mlb_launch_count() - l0
        dev_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev_ms, wall = t[0].item(), t[1].item() / 1e3
        barrier()
        return dev_ms / 1e3, wall, launches, clocks

    # ---- headline: device-resident
    dev_s, _, launches, clocks = timed(step_device, args.steps, args.warmup, sample_clocks=True)
    pts_per_step = world * n_items * K * K
    value = pts_per_step * args.steps / dev_s

    # ---- e2e: pinned host -> H2D -> kernels -> D2H every step
    e2e_steps = max(3, min(args.steps, 10))
    _, e2e_wall, _, _ = timed(step_host, e2e_steps, 2)
    e2e_value = pts_per_step * e2e_steps / e2e_wall
    h2d = world * n_items * plans[0].h2d_bytes          # whole job, all ranks
    d2h = world * n_items * plans[0].d2h_bytes

    # ---- the other formulations on the same workload (device-resident, fewer steps), for context
    paths = {plans[0].method: value}
    if world == 1 and not args.no_paths:
        for m in ("fft", "fold", "tc", "dense"):
            if m in paths:
                continue
            try:
                alt = [FarfieldPlan((M, M), p.dxp, p.dyp, p.wavelength, p.n_glass, stride=s, method=m) for p in plans]
                items_alt = [t.item for t in sharded.tiles]
            except ValueError:
                continue

            def step_alt(alt=alt):
                for it, plan in zip(items_alt, alt):
                    plan.run(fields_of(it))
            t_alt, _, _, _ = timed(step_alt, 3, 2)
            paths[m] = pts_per_step * 3 / t_alt
            del alt
            torch.cuda.empty_cache()

    # ---- per-kernel timing of one item for the roofline (CUDA events on the launching stream)
    def kernel_time(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    # every kernel of one item, timed alone; rotate over the batch items so that the inputs
    # (n_items x 32 M^2 bytes) exceed L2 between launches of the same kernel
    all_steps = [plan.steps(fields_of(t.item)) for t, plan in zip(sharded.tiles, plans)]
    for t, plan in zip(sharded.tiles, plans):
        plan.run(fields_of(t.item))
    kernels = {}
    rr = [0]
    for k, (name, _fn, nbytes, flops) in enumerate(all_steps[0]):
        def one(k=k):
            rr[0] = (rr[0] + 1) % len(all_steps)
            all_steps[rr[0]][k][1]()
        kernels[name] = dict(seconds=kernel_time(one), bytes=nbytes, flops=flops)
    p0 = plans[0]
    dom = max(kernels, key=lambda k: kernels[k]["seconds"])
    kd = kernels[dom]
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12       # nominal fp32 FMA pipe, TFLOP/s at max clock
    if not dom.startswith("cgemm"):
        ach = kd["bytes"] / kd["seconds"] / 1e9
        roof = dict(kernel=dom, bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=ach / peaks["hbm_gbs"], traffic=None, peak_source=peaks["source"])
    else:
        ach = kd["flops"] / kd["seconds"] / 1e12
        roof = dict(kernel=dom, bound="fp32-fma (SIMT; not a tensor-core kernel)", achieved=ach, peak=fp32_peak,
                    unit="TFLOP/s", frac=ach / fp32_peak, traffic=None,
                    peak_source="nominal 148 SM x 128 FMA/clk x 2 x 1.965 GHz; bf16 tensor peak %s TF/s for context"
                                % peaks["bf16_tflops"])
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        t = json.load(open(tr)).get(args.workload, {}).get(dom)
        if t:
            roof["traffic"] = t["dram_bytes_per_launch"]
            roof["traffic_source"] = t["source"]
    step_kernel_s = sum(k["seconds"] for k in kernels.values())
    roof["share_of_step"] = kd["seconds"] / step_kernel_s
    if "fold" in kernels:
        a = kernels["fold"]["bytes"] / kernels["fold"]["seconds"] / 1e9
        roof["fold_hbm"] = dict(achieved=a, peak=peaks["hbm_gbs"], unit="GB/s", frac=a / peaks["hbm_gbs"])

    # ---- BASELINE configs[1] (2048^2 -> 512^2, TE+TM) on the same code path, for the record
    other = None
    if world == 1 and not args.no_paths and args.workload == "cfg3":
        w2 = WORKLOADS["cfg2"]
        M2, K2 = w2["M"], w2["M"] // w2["stride"]
        f2, p2 = [], []
        for i, (wl, ng, rot) in enumerate(w2["items"]):
            Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M2, 2000 + i, wl, ng, rotate=rot)
            f2.append([torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)])
            p2.append(FarfieldPlan((M2, M2), float(x[1] - x[0]), float(x[1] - x[0]), wl, ng, stride=w2["stride"],
                                   method=args.method))

        def step_cfg2():
            for pl, fl in zip(p2, f2):
                pl.run(fl)
        t2, _, _, _ = timed(step_cfg2, args.steps, args.warmup)
        other = {"cfg2": {"workload": w2["name"], "method": p2[0].method,
                          "value": len(p2) * K2 * K2 * args.steps / t2, "ms_per_step": t2 / args.steps * 1e3,
                          "note": "268 MB of inputs per step: larger than L2"}}
        del f2, p2
        # the reference's own usage: a good_fft_number() grid (3375 = 3^3 5^3), ALL FFT bins
        M3 = 3375
        wl3, ng3 = 580e-9, apertures.N_GLASS[580]
        Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M3, 2100, wl3, ng3)
        f3 = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
        p3 = FarfieldPlan((M3, M3), float(x[1] - x[0]), float(x[1] - x[0]), wl3, ng3, stride=1, method=args.method)
        t3, _, _, _ = timed(lambda: p3.run(f3), 5, 2)
        other["ref_default_grid"] = {"workload": "3375x3375 aperture (good_fft_number size) -> all 3375x3375 FFT bins, 580 nm",
                                     "method": p3.method, "value": M3 * M3 * 5 / t3, "ms_per_step": t3 / 5 * 1e3,
                                     "note": "big-radix mixed engine (3 x radix-15 stages in registers; columns as 15 x 225 in two passes)"}
        del f3, p3
        torch.cuda.empty_cache()

    nf = None
    if rank == 0 and world == 1 and not args.no_nearfield:
        del dev_fields, pinned
        torch.cuda.empty_cache()
        nf = bench_nearfield(args.nearfield_m, torch)

    sweep = None
    if rank == 0 and world == 1 and not args.no_paths:
        sweep = bench_sweep(torch)

    if rank == 0:
        line = {
            "metric": "far-field points/sec (NF->FF)", "value": value, "unit": "far-field points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (complex64 fields, fp32 accumulate, f64 twiddle phases/epilogue)",
            "data": "synthetic",
            "config": {"workload": w["name"], "method": p0.method, "batch_items_per_gpu": n_items,
                       "aperture": [M, M], "far_field": [K, K], "l2": "inputs larger than L2 (%.0f MB per step)"
                       % (n_items * 32 * M * M / 1e6), "parallelism": "far-field tiles sharded, %d rank(s)%s" % (
                           world, "" if world == 1 else ", tile exchange: " + {"p2p": "peer-to-peer pulls over NVLink (copy "
                           "engines, symmetric memory)", "nccl": "NCCL all-gather"}.get(sharded._gather, sharded._gather))},
            "e2e": {"value": e2e_value, "unit": "far-field points/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": {k: dict(ms=v["seconds"] * 1e3, gbs=v["bytes"] / v["seconds"] / 1e9,
                                tflops=v["flops"] / v["seconds"] / 1e12) for k, v in kernels.items()},
            "paths_points_per_s": paths,
            "other_workloads": other,
            "nearfield_assembly": nf,
            "fom_sweep": sweep,
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """Libraries (NCCL prints its version banner, torchrun its OMP note) must not pollute the ONE JSON line:
    everything written to fd 1 from here on goes to stderr; the JSON line is written to the saved fd."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    os.write(_STDOUT_FD, data)


_STDOUT_FD = 1


def main():
    global _STDOUT_FD
    _STDOUT_FD = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--method", default="auto", choices=["auto", "dense", "fold", "fft", "tc"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-paths", action="store_true", help="skip timing the alternative formulations")
    ap.add_argument("--no-nearfield", action="store_true", help="skip the aperture-assembly (hot path B) section")
    ap.add_argument("--nearfield-m", type=int, default=4096, help="aperture size of the hot path B section")
    ap.add_argument("--quick-cpu", action="store_true", help="tiny cpu_baseline sample (debug)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
