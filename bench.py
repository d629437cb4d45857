#!/usr/bin/env python
"""bench.py -- far-field points/s of the NF->FF hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4]

One "step" = one pass of the hot path over one batch: every batch item (wavelength /
polarisation) of the workload goes aperture fields -> far-field power map P.

  cfg3 (default; the configuration the north-star target is quoted on):
       4096^2 aperture -> 1024^2 far field (every 4th fftshifted FFT bin), 3 wavelengths
  cfg2 2048^2 aperture -> 512^2 far field, 532 nm, TE+TM (2 items)
  cfg4 ONE 8192^2 aperture assembled from a synthetic NA 0.94 lens (hot path B) -> 2048^2 far field, the whole
       step spread over the N GPUs (strong scaling; metalens_b200/slab.py).  Also reported as the `cfg4` block of
       the default line at every N.

value  = far-field points/s with the aperture fields already resident in HBM
e2e    = same metric through FarfieldPlan.run_host(): pinned host fields -> H2D ->
         kernels -> D2H of P, every step
N > 1  : one process per GPU (torchrun), weak scaling -- every rank owns a full batch of its
         own apertures (far-field tiles of different sources); one all-gather of the P tiles per
         step is inside the timed region (asynchronous, double-buffered: it overlaps the kernels of
         the next step).  Every rank pushes its tiles into the peers' result buffers over NVLink with
         the library's own kernel (mlb_peer_allgather; torch symmetric memory only maps the buffers);
         the C-ABI's NCCL wrapper is the fallback where peer memory cannot be mapped.
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
    "cfg4": dict(M=8192, stride=4, items=[(580e-9, 1.459, False)],
                 name="cfg4: NA 0.94 synthetic lens (3 GratingCollections + HexGridSet), ONE 8192x8192 aperture assembled "
                      "on the GPUs -> 2048x2048 far field (every 4th FFT bin), 580 nm, whole step sharded over the ranks"),
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
def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ReferenceRunner:
    """The reference's own CPU implementation of hot path A on the workload's apertures.

    kind "reference": the UNMODIFIED nearfield_farfield.farfield_from_nearfield from baseline/_ref (installed by
    __graft_entry__.build() from /root/reference; baseline/install_ref.py), fed the way its docstring prescribes
    (nearfield_farfield.py:18-20): fft2(fftshift(E)) of the four complex128 fields, computed by the caller with
    numpy.fft.  Host threads: the batch items run concurrently and each item's four FFTs run concurrently (numpy
    releases the GIL); farfield_from_nearfield itself is single-threaded numpy, as shipped.
    kind "port": oracle/farfield_oracle.py (same algorithm, threaded chunk loop) when baseline/_ref is absent."""

    def __init__(self, workload, sample_M=None):
        w = WORKLOADS[workload]
        self.w = w
        self.M = w["M"] if sample_M is None else sample_M
        self.K = self.M // w["stride"]
        self.cores = _host_cores()
        sys.path.insert(0, ROOT)
        from baseline import install_ref
        self.ref = install_ref.load()
        self.kind = "reference" if self.ref is not None else "port"
        self.inputs = []
        for i, (wl, ng, rot) in enumerate(w["items"]):
            Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(self.M, 100 + i, wl, ng, rotate=rot)
            # the reference's build_nearfield hands over complex128 fields
            self.inputs.append(([a.astype(np.complex128) for a in (Ex, Ey, Hx, Hy)], x, y, wl, ng))

    def _one(self, item):
        import contextlib
        import io
        from concurrent.futures import ThreadPoolExecutor
        fields, x, y, wl, ng = self.inputs[item]
        if self.kind == "port":
            from oracle import farfield_oracle as fo
            per_item = max(1, self.cores // len(self.inputs))
            return float(fo.farfield_reference_path_threads(*fields, x, y, wl, ng, workers=per_item)[1])
        with ThreadPoolExecutor(4) as pool:
            F = list(pool.map(lambda a: np.fft.fft2(np.fft.fftshift(a)), fields))             # :18-20
        with contextlib.redirect_stdout(io.StringIO()):                                       # progress prints, :53
            out = self.ref["nearfield_farfield"].farfield_from_nearfield(*F, x, y, wl, ng)   # :14
        return float(out[1])

    def step(self, items):
        """One pass over `items` (indices into the batch), items concurrently.  Returns seconds."""
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        if len(items) > 1 and self.cores > 1:
            with ThreadPoolExecutor(len(items)) as pool:
                list(pool.map(self._one, items))
        else:
            for it in items:
                self._one(it)
        return time.perf_counter() - t0

    def threads_used(self, n_items):
        if self.kind == "port":
            return min(self.cores, 64)
        return min(self.cores, 4 * n_items)

    def describe(self, items_per_step, t):
        return ("%d of %d batch items per step, aperture %dx%d -> %dx%d requested bins (the numpy path always computes "
                "all %d^2 bins: %.3g bins/s); %s, float64, %d host thread(s) (items and the four caller-side FFTs "
                "concurrent), %.2f s per step"
                % (items_per_step, len(self.inputs), self.M, self.M, self.K, self.K, self.M,
                   items_per_step * self.M ** 2 / t,
                   "UNMODIFIED reference farfield_from_nearfield (baseline/_ref) + numpy.fft.fft2(fftshift(.))"
                   if self.kind == "reference" else "oracle/farfield_oracle.py port of nearfield_farfield.py",
                   self.threads_used(items_per_step), t))


def cpu_baseline_leg(workload, quick=False):
    """cpu_baseline of our arm: ONE batch item of the workload at full aperture size (about 5-15 s of CPU work),
    through the same runner as --impl reference."""
    r = ReferenceRunner(workload, sample_M=512 if quick else None)
    t = r.step([0])
    return dict(value=r.K * r.K / t, unit="far-field points/s", cores=r.threads_used(1), kind=r.kind,
                sample=r.describe(1, t))


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path at the workload's FULL size, honouring
    --steps / --warmup.  One full-size step is measured first; if (steps + warmup) such steps do not fit the time
    budget, every step processes ONE batch item (rotating through the wavelengths: same aperture, same per-point
    work) and, if even that does not fit, the warm-up shrinks (never below 1) -- the line says what was run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload if args.workload in ("cfg3", "cfg2") else "cfg3"
    r = ReferenceRunner(workload)
    n = len(r.inputs)
    budget = args.ref_budget
    t_start = time.perf_counter()
    rot = 0

    def next_items():
        nonlocal rot
        if per_step == n:
            return list(range(n))
        rot = (rot + 1) % n
        return [rot]
    per_step = n
    t_full = r.step(next_items())                           # warm-up step 1: the full batch
    steps, warmup, warmed = args.steps, max(args.warmup, 1), 1
    if (steps + warmup - 1) * t_full > budget and n > 1:
        per_step = 1
        t_one = r.step(next_items())                        # warm-up step 2: what a step will be from now on
        warmed = 2
        while warmup > warmed and (steps + warmup - warmed) * t_one > budget:
            warmup -= 1
        warmup = max(warmup, warmed)
    for _ in range(warmup - warmed):
        r.step(next_items())
    times = []
    for _ in range(steps):
        times.append(r.step(next_items()))
        if time.perf_counter() - t_start > 3 * budget:      # hard stop: never run away
            break
    t = float(np.mean(times))
    value = per_step * r.K * r.K / t
    base = dict(value=value, unit="far-field points/s", cores=r.threads_used(per_step), kind=r.kind,
                sample=r.describe(per_step, t))
    line = {
        "impl": "reference", "metric": "far-field points/sec (NF->FF)", "value": value,
        "unit": "far-field points/s", "n_gpus": args.gpus, "steps": len(times), "warmup": warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": common_config(WORKLOADS[workload], r.M, r.K, n),
        "details": {"batch_items_per_step": per_step, "sample": base["sample"], "budget_s": budget,
                    "requested_steps": args.steps, "requested_warmup": args.warmup},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "far-field points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------- hot path B
def nearfield_cpu_baseline(M, spec, x, periph, center_rows, hgs, cpu_cols=384):
    """Hot path B on the host, on a strip of the same grid (off-centre, so centre and rings are both represented).
    kind "reference": the UNMODIFIED nearfield.build_nearfield from baseline/_ref, fed a library built with the
    reference's own classes (its interpolators are scipy objects) and the same design rows; single-threaded as
    shipped.  kind "port": oracle/nearfield_oracle.py when baseline/_ref is absent (or the reference run fails)."""
    import contextlib
    import io
    import synth_lens
    wl, f = 580e-9, spec["source_distance"]
    j0 = M // 2 + M // 8
    ys = x[j0:j0 + cpu_cols]
    sys.path.insert(0, ROOT)
    try:
        from baseline import install_ref
        ref = install_ref.load()
        if ref is not None:
            from metalens_b200.design import make_design
            collections, r_hgs = synth_lens.make_library(ref["grating"], ref["lens_center"], spec)
            r_periph, r_center, _ = make_design(collections, f, spec["radius"], r_hgs)      # host layout, same rows
            assert len(r_center) == len(center_rows)
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):                                  # per-order progress prints
                ref["nearfield"].build_nearfield(0.0, 0.0, -f, "x", wl, r_periph, r_center, r_hgs, x_pts=x, y_pts=ys)
            tc = time.perf_counter() - t0
            return dict(value=M * cpu_cols / tc, unit="aperture samples/s", cores=1, kind="reference",
                        sample="%d x %d strip of the same grid, UNMODIFIED reference nearfield.build_nearfield "
                               "(baseline/_ref; scipy interpolators + cKDTree), 1 thread as shipped, %.1f s" % (M, cpu_cols, tc))
    except Exception as e:                                          # noqa: BLE001 -- fall back to the port, say why
        print("bench: reference build_nearfield unavailable (%s: %s); timing the oracle port" % (type(e).__name__, str(e)[:200]),
              file=sys.stderr)
    from oracle import nearfield_oracle as no
    t0 = time.perf_counter()
    no.build_nearfield(0.0, 0.0, -f, "x", wl, periph, center_rows, hgs, x_pts=x, y_pts=ys)
    tc = time.perf_counter() - t0
    return dict(value=M * cpu_cols / tc, unit="aperture samples/s", cores=1, kind="port",
                sample="%d x %d strip of the same grid, oracle/nearfield_oracle.py (numpy + scipy "
                       "cKDTree), 1 thread, %.1f s" % (M, cpu_cols, tc))


def bench_nearfield(M, torch, peaks, cpu_cols=384):
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
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    make_design(collections, f, spec["radius"], hgs, device="cuda")           # first call: kernel loading
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    periph, center, _ = make_design(collections, f, spec["radius"], hgs, device="cuda")   # hex centre laid out on the GPU (N2)
    torch.cuda.synchronize()
    t_design = time.perf_counter() - t0
    t0 = time.perf_counter()
    plan = NearfieldPlan(wl, periph, center, hgs)
    torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0
    x = np.linspace(-R, R, M)
    out = torch.zeros((4, M, M), dtype=torch.complex64, device="cuda")
    for _ in range(2):
        plan.run(0.0, 0.0, -f, "x", x, x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        plan.run(0.0, 0.0, -f, "x", x, x, out=out, check=False)      # asynchronous launches; bounds checked below
    e1.record()
    torch.cuda.synchronize()
    plan.check_violation()
    t = e0.elapsed_time(e1) / reps * 1e-3
    res = dict(aperture=[M, M], rings=int(len(periph["r_min_list"])), hex_cells=int(len(center)),
               ms=t * 1e3, samples_per_s=M * M / t, write_gbs=32.0 * M * M / t / 1e9,
               roofline=dict(kernel="nearfield_kernel", bound="hbm", achieved=32.0 * M * M / t / 1e9, peak=peaks["hbm_gbs"],
                             unit="GB/s", frac=32.0 * M * M / t / 1e9 / peaks["hbm_gbs"], traffic=None,
                             algorithmic_bytes_per_launch=32 * M * M,
                             note="SURVEY 8(d): bytes_B = 32*M^2 written; the kernel is bound by instruction issue "
                                  "(float64 geometry), see profiles/"),
               note="one fused kernel launch + the incident-power sum per call, bounds violation checked after the timed "
                    "loop; algorithmic bytes = 32*M^2 written (4 complex64 fields)", design_seconds=t_design,
               plan_seconds=t_plan, design_note="make_design with the hex centre on the device (mlb_hex_count / mlb_hex_fill) and "
                                                "NearfieldPlan (device binning, table packs, ring slices), wall clock")
    res["cpu_baseline"] = nearfield_cpu_baseline(M, spec, x, periph, center.cpu().numpy(), hgs, cpu_cols)
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


# ----------------------------------------------------------------------------- cfg4: one aperture over N GPUs
def cfg4_lens(M, wl=580e-9):
    """synthetic NA ~0.94 lens filling an M x M grid (lambda/2.2 sampling)."""
    import synth_lens
    from metalens_b200 import grating, lens_center
    from metalens_b200.design import make_design
    R = M * (wl / 2.2) / 2
    f = R / math.tan(math.asin(0.94))
    spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1), (45.0, 71.0, 300e-9, 2.3)],
                source_distance=f, radius=R * 0.999)
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph, center, _ = make_design(collections, f, spec["radius"], hgs, device="cuda")
    return periph, center, hgs, R, f


def bench_cfg4(torch, dist, rank, world, steps, warmup, M=8192, stride=4, check=True, sample_clocks=None):
    """BASELINE config 4 as ONE job over `world` GPUs (strong scaling): every rank assembles its rows of the aperture
    (hot path B), folds + row-transforms them while scattering the column slabs to their owners over NVLink, column
    pass + power on its slab, one pushed all-gather of P (metalens_b200/slab.py).  Timed: K steps of
    [assembly -> far field complete on every rank], CUDA events, max over ranks.  On rank 0 the result is compared
    BIT FOR BIT with the single-GPU pipeline (full assembly -> FarfieldPlan) after the timed region."""
    from metalens_b200 import _lib
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.nearfield import NearfieldPlan
    from metalens_b200.peer import SymmetricPeers, VirtualPeers
    from metalens_b200.slab import SlabFarfield, assemble_slab
    lib = _lib.load()
    wl = 580e-9
    t0 = time.perf_counter()
    periph, center, hgs, R, f = cfg4_lens(M, wl)
    t_design = time.perf_counter() - t0
    nf = NearfieldPlan(wl, periph, center, hgs)
    x = np.linspace(-R, R, M)
    d = float(x[1] - x[0])
    peers = SymmetricPeers() if world > 1 else VirtualPeers(1).view(0)
    slab = SlabFarfield((M, M), d, d, wl, nf.n_glass, stride, peers)
    K = slab.K1
    out = torch.zeros((4, slab.x_rows.size, M), dtype=torch.complex64, device="cuda")
    fields = [out[i] for i in range(4)]
    src = (0.0, 0.0, -f)
    power = [None]

    def assemble():
        power[0] = assemble_slab(nf, slab, src, "x", x, x, out=out, check=False)[1]

    def farfield():
        return slab.run(fields)

    def step():
        assemble()
        return farfield()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n, w):
        for _ in range(w):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.mlb_launch_count()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        launches = lib.mlb_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        barrier()
        return ms, launches

    sampler = sample_clocks() if sample_clocks else None
    ms_step, launches = timed(step, steps, warmup)
    clocks = sampler.stop() if sampler else None
    ms_nf, _ = timed(assemble, max(3, steps // 2), 2)
    ms_ff, _ = timed(farfield, max(3, steps // 2), 2)
    nf.check_violation()
    slab.chan.check()
    P, total = step()
    torch.cuda.synchronize()
    p_in = power[0].clone()
    if world > 1:
        dist.all_reduce(p_in, op=dist.ReduceOp.SUM)
    res = dict(workload=WORKLOADS["cfg4"]["name"], aperture=[M, M], far_field=[K, K], n_gpus=world,
               rings=int(len(periph["r_min_list"])), hex_cells=int(center.shape[0]), design_seconds=t_design,
               ms_per_step=ms_step, ms_assembly=ms_nf, ms_farfield=ms_ff, steps=steps, warmup=warmup,
               far_field_points_per_s=K * K / ms_step * 1e3, aperture_samples_per_s=M * M / ms_step * 1e3,
               assembly_write_gbs=32.0 * M * M / world / ms_nf / 1e6,
               nvlink_bytes_per_rank=dict(all_to_all=32 * K * K * (world - 1) // world ** 2,
                                          gather=4 * K * K * (world - 1) // world),
               total_P=float(total), incident_power=float(p_in), gpu_launches_per_step=launches / steps,
               scaling="strong", exchange="peer stores over NVLink (mlb_fft_rows_scatter, mlb_peer_barrier, "
                                          "mlb_peer_allgather)" if world > 1 else "single GPU",
               note="total_P is the Riemann sum of P over every 4th FFT bin: a collimator radiates into a spot about one "
                    "bin wide, which that coarse quadrature over-weights (tests/test_full_configs_gpu.py::"
                    "test_strided_total_is_a_subsampled_sum); P itself is exact at every sampled bin")
    if clocks is not None:
        res["clocks"] = clocks
    if check and rank == 0:
        # the same lens on ONE GPU: full assembly, single-GPU far field with the same kernels
        full = torch.zeros((4, M, M), dtype=torch.complex64, device="cuda")
        nf.run(src[0], src[1], src[2], "x", x, x, out=full)
        plan = FarfieldPlan((M, M), d, d, wl, nf.n_glass, stride=stride, method="fft", fuse_power="always")
        old = lib.mlb_get_option(b"rows_engine")
        lib.mlb_set_option(b"rows_engine", 0)
        P1, total1 = plan.run([full[i] for i in range(4)])
        lib.mlb_set_option(b"rows_engine", old)
        torch.cuda.synchronize()
        same = bool(((P == P1) | (torch.isnan(P) & torch.isnan(P1))).all())
        res["bit_identical_to_single_gpu"] = same and float(total) == float(total1)
        assert res["bit_identical_to_single_gpu"], "cfg4: sharded far field differs from the single-GPU result"
        del full, plan, P1
        torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------- GPU arm
def common_config(w, M, K, n_items):
    """`config` keys shared by both arms (ours / --impl reference), so that the two lines name the same job."""
    return {"workload": w["name"], "aperture": [M, M], "far_field": [K, K], "batch_items_per_gpu": n_items,
            "l2": "inputs larger than L2 (%.0f MB of complex64 fields per step and GPU)" % (n_items * 32 * M * M / 1e6)}


def ours(args):
    import torch
    import torch.distributed as dist
    from metalens_b200 import _lib, hostmem
    from metalens_b200.farfield import FarfieldPlan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first (rank 0, N=1 only), before the GPU work so that the host cores are idle
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_leg(args.workload if args.workload in ("cfg2", "cfg3") else "cfg3", quick=args.quick_cpu)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    peaks = measured_peaks()

    if args.workload == "cfg4":
        c4 = bench_cfg4(torch, dist, rank, world, args.steps, args.warmup,
                        sample_clocks=lambda: _started(ClockSampler(local)))
        if rank == 0:
            K = c4["far_field"][0]
            line = {
                "metric": "far-field points/sec (NF->FF)", "value": c4["far_field_points_per_s"],
                "unit": "far-field points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": c4["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32 (complex64 fields, fp32 accumulate, f64 geometry/phases)",
                "data": "synthetic",
                "config": {"workload": c4["workload"], "aperture": c4["aperture"], "far_field": c4["far_field"],
                           "l2": "the step writes and re-reads a 2.1 GB aperture: larger than L2",
                           "parallelism": "one aperture over %d rank(s): rows of the assembly + fold/row FFT per rank, "
                                          "all-to-all fused into the row pass (peer stores), column slabs, one pushed "
                                          "all-gather of P" % world},
                "e2e": {"value": c4["far_field_points_per_s"], "unit": "far-field points/s",
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "the step starts from the lens description (device-resident tables): no host fields exist"},
                "gpu_launches": int(round(c4["gpu_launches_per_step"] * args.steps)), "clocks": c4.pop("clocks", None),
                "roofline": dict(kernel="nearfield_kernel", bound="hbm", achieved=c4["assembly_write_gbs"],
                                 peak=peaks["hbm_gbs"], unit="GB/s", frac=c4["assembly_write_gbs"] / peaks["hbm_gbs"],
                                 traffic=None, peak_source=peaks["source"],
                                 note="algorithmic bytes = 32*M^2/G written per rank; the kernel is instruction-bound"),
                "cfg4": c4, "cpu_baseline": cpu,
            }
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    w = WORKLOADS[args.workload]
    M, s = w["M"], w["stride"]
    K = M // s
    n_items = len(w["items"])

    # far-field tiles over the ranks (metalens_b200/sharding.py): weak scaling -> world x n_items batch
    # items, whole items per rank, so each rank synthesises and owns only its own apertures
    from metalens_b200.sharding import ShardedFarfield
    n_global = n_items * world

    def make_plan(item, r0, r1):
        wl, ng, rot = w["items"][item % n_items]
        d = apertures.grid(M, wl)[0]
        d = float(d[1] - d[0])
        rows = None if (r0, r1) == (0, K) else (r0, r1)
        return FarfieldPlan((M, M), d, d, wl, ng, stride=s, method=args.method if rows is None else "fold", rows=rows)

    sharded = ShardedFarfield(n_global, K, make_plan, rank=rank, world=world, gather=args.gather,
                              push_ctas=args.push_ctas)
    plans = sharded.plans
    pinned, dev_fields = {}, {}
    numa_cpus = (hostmem.gpu_local_cpus(local) or set()) & os.sched_getaffinity(0)
    if numa_cpus == os.sched_getaffinity(0):
        numa_cpus = set()
    for item in sharded.items_needed:
        wl, ng, rot = w["items"][item % n_items]
        Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 1000 + item, wl, ng, rotate=rot)
        pin = hostmem.pinned_empty((4, M, M), torch.complex64, local)      # pages next to this rank's GPU
        for f, a in enumerate((Ex, Ey, Hx, Hy)):
            pin[f].copy_(torch.from_numpy(a))
        pinned[item] = pin
        dev_fields[item] = pin.cuda()
        del Ex, Ey, Hx, Hy

    def fields_of(item):
        return [dev_fields[item][f] for f in range(4)]

    def step_eager():
        # local tiles + the one all-gather, which overlaps the next step's kernels (double-buffered)
        return sharded.run(fields_of, overlap=True)

    # Default: eager launches.  They are asynchronous and the host (~0.2 ms per step) stays ahead of the device, so the
    # two-stream pipeline runs ACROSS steps: the column tail of a step's last item executes under the next step's
    # first row pass.  Replaying the step as one CUDA graph (--graph; ShardedFarfield.capture/replay) removes the host
    # cost but serialises consecutive steps at the graph boundary, which exposes that tail: measured slower for
    # cfg3, reported as details.graph_replay_ms_per_step.
    step_device = step_eager

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
        sharded.finish()
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
        sharded.finish()                          # outstanding exchanges: every peer's tiles have landed
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
    graph_ms, graphed = None, False
    if args.graph:
        try:
            sharded.capture(fields_of)
            graphed = True
        except Exception as e:                                 # noqa: BLE001 -- report and carry on
            print("bench: CUDA-graph capture of the step unavailable (%s: %s)" % (type(e).__name__, str(e)[:200]),
                  file=sys.stderr)
        if world > 1:                                          # every rank must take the same path
            flag = torch.tensor([1 if graphed else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            graphed = bool(flag.item())
        if graphed:
            g_s, _, _, _ = timed(sharded.replay, args.steps, 3)
            graph_ms = g_s / args.steps * 1e3
            if args.graph and graph_ms < dev_s / args.steps * 1e3:
                dev_s = g_s                                    # --graph: the replayed step is the headline if it is faster
            else:
                graphed = False
    sharded.check()
    sharded_gather = sharded._gather
    pts_per_step = world * n_items * K * K
    value = pts_per_step * args.steps / dev_s

    # ---- parity inside the bench (rank 0): the gathered tiles of the LAST step against single-GPU plans
    gathered_ok = None
    if world > 1:
        P_all, _ = step_eager()
        sharded.finish()
        torch.cuda.synchronize()
        mine = sharded.items_needed
        ok = True
        for it, plan in zip(mine, plans):
            Pl, _t = plan.run(fields_of(it))
            ok = ok and bool(((P_all[it] == Pl) | (torch.isnan(P_all[it]) & torch.isnan(Pl))).all())
        # foreign items: same apertures (seeded by item number modulo the batch) are NOT on this rank; check instead that
        # every rank received identical bytes (checksum all-reduce MIN == MAX)
        cs = torch.nan_to_num(P_all.double(), nan=0.0).sum(dim=(1, 2))
        lo, hi = cs.clone(), cs.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ok = ok and bool(torch.equal(lo, hi))
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gathered_ok = bool(flag.item())
        assert gathered_ok, "gathered far-field tiles differ from the single-GPU result / between ranks"

    # ---- e2e: pinned host -> H2D -> kernels -> D2H every step
    e2e_steps = max(3, min(args.steps, 10))
    _, e2e_wall, _, _ = timed(step_host, e2e_steps, 2)
    e2e_value = pts_per_step * e2e_steps / e2e_wall
    h2d = world * n_items * plans[0].h2d_bytes          # whole job, all ranks
    d2h = world * n_items * plans[0].d2h_bytes

    # ---- the other formulations on the same workload (device-resident, fewer steps), for context
    paths = {plans[0].method: value}
    if world == 1 and not args.no_paths:
        for m in ("fft", "fold", "czt", "tc", "dense"):
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
    # algorithmic bytes of NF->FF per item = 32*M^2 (four complex64 fields read once) + 4*K^2 (P written
    # once).  The kernel that touches the aperture is charged the 32*M^2 it must read; the K x K intermediate it also
    # writes is not algorithmic
    algo = {"fold_fft_rows": 32 * M * M, "fft_rows": 32 * M * M, "fold": 32 * M * M}
    if not dom.startswith("cgemm") and not dom.startswith("tc_"):
        ab = algo.get(dom, kd["bytes"])
        ach = ab / kd["seconds"] / 1e9
        roof = dict(kernel=dom, bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=ach / peaks["hbm_gbs"], traffic=None, peak_source=peaks["source"],
                    algorithmic_bytes_per_launch=ab, us_per_launch=kd["seconds"] * 1e6)
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
    # the whole step against the floor of its algorithmic bytes at the measured HBM peak
    step_bytes = n_items * (32 * M * M + 4 * K * K)
    roof["step"] = dict(algorithmic_bytes=step_bytes, ms=dev_s / args.steps * 1e3,
                        achieved=step_bytes / (dev_s / args.steps) / 1e9,
                        frac=step_bytes / (dev_s / args.steps) / 1e9 / peaks["hbm_gbs"])
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

        # two small items per step: launch-bound when issued eagerly, so the step is replayed as CUDA graphs
        # (ShardedFarfield.capture / replay, the API a sweep would use)
        sh2 = ShardedFarfield(len(p2), K2, lambda item, r0, r1: p2[item], rank=0, world=1)
        cfg2_mode = "eager"
        try:
            sh2.capture(lambda item: f2[item])
            step_cfg2 = sh2.replay
            cfg2_mode = "CUDA-graph replay"
        except Exception:                                      # noqa: BLE001
            def step_cfg2():
                for pl, fl in zip(p2, f2):
                    pl.run(fl)
        t2, _, _, _ = timed(step_cfg2, args.steps, args.warmup)
        sh2.finish()
        other = {"cfg2": {"workload": w2["name"], "method": p2[0].method, "step": cfg2_mode,
                          "value": len(p2) * K2 * K2 * args.steps / t2, "ms_per_step": t2 / args.steps * 1e3,
                          "note": "268 MB of inputs per step: larger than L2"}}
        del sh2
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
                                     "note": "big-radix mixed engine (rows: 3 x radix-15 stages in registers, kernel compiled for the plan; columns as 15 x 225 in two passes)"}
        # the strict drop-in (reference signature: host complex128 FFT'd fields in, host float64 P out; only the
        # epilogue runs on the GPU) on the same grid, wall clock
        from metalens_b200.farfield import farfield_from_nearfield
        F = [np.fft.fft2(np.fft.fftshift(a.astype(np.complex128))) for a in (Ex, Ey, Hx, Hy)]
        farfield_from_nearfield(*F, x, y, wl3, ng3)
        t0 = time.perf_counter()
        farfield_from_nearfield(*F, x, y, wl3, ng3)
        td = time.perf_counter() - t0
        other["dropin_farfield_from_nearfield"] = {
            "workload": "reference signature, 3375x3375 FFT'd complex128 host arrays -> float64 host P",
            "seconds": td, "value": M3 * M3 / td,
            "note": "pageable H2D of 4 x 182 MB + epilogue + D2H; the caller-side numpy FFTs are not included"}
        del f3, p3, F
        torch.cuda.empty_cache()

    nf = None
    if rank == 0 and world == 1 and not args.no_nearfield:
        nf = bench_nearfield(args.nearfield_m, torch, peaks)

    sweep = None
    if rank == 0 and world == 1 and not args.no_paths:
        sweep = bench_sweep(torch)

    # ---- BASELINE config 4 as ONE job over all ranks (strong scaling), every N
    c4 = None
    if not args.no_cfg4:
        del dev_fields, pinned, sharded, plans, all_steps
        torch.cuda.empty_cache()
        c4 = bench_cfg4(torch, dist, rank, world, max(5, args.steps // 2), 3)

    if rank == 0:
        cfg = common_config(w, M, K, n_items)
        line = {
            "metric": "far-field points/sec (NF->FF)", "value": value, "unit": "far-field points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (complex64 fields, fp32 accumulate, f64 twiddle phases/epilogue)",
            "data": "synthetic",
            "config": cfg,
            "details": {"method": p0.method, "step": "one CUDA-graph launch per step" if graphed else "eager launches",
                        "graph_replay_ms_per_step": graph_ms,
                        "parallelism": "far-field tiles sharded, %d rank(s)%s" % (
                            world, "" if world == 1 else ", tile exchange: " + {
                                "push": "every rank pushes its tiles into the peers' result buffers over NVLink "
                                        "(mlb_peer_allgather, peer-mapped symmetric memory)",
                                "p2p": "peer-to-peer pulls over NVLink (copy engines, symmetric memory)",
                                "nccl": "NCCL all-gather (mlb_allgather_P)"}.get(sharded_gather, sharded_gather)),
                        "gathered_tiles_bit_identical": gathered_ok},
            "e2e": {"value": e2e_value, "unit": "far-field points/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "h2d_gbs": h2d / world * e2e_steps / e2e_wall / 1e9,
                    "host_buffers": "pinned on the CPUs local to the GPU (%s)" % (
                        "rank 0: %d of %d CPUs" % (len(numa_cpus), len(os.sched_getaffinity(0))) if numa_cpus
                        else "rank 0: topology unknown or a single node, plain pinned memory"),
                    "note": "PCIe-bound: 32*M^2 bytes of host fields per item against ~0.1 ms of kernels"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": {k: dict(ms=v["seconds"] * 1e3, gbs=v["bytes"] / v["seconds"] / 1e9,
                                tflops=v["flops"] / v["seconds"] / 1e12) for k, v in kernels.items()},
            "paths_points_per_s": paths,
            "other_workloads": other,
            "nearfield_assembly": nf,
            "fom_sweep": sweep,
            "cfg4": c4,
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _started(sampler):
    sampler.start()
    time.sleep(0.25)
    return sampler


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
    ap.add_argument("--no-cfg4", action="store_true", help="skip the cfg4 (one aperture over all ranks) block")
    ap.add_argument("--graph", action="store_true", help="also time the step replayed as CUDA graphs (ShardedFarfield."
                    "capture/replay); it becomes the headline when it is faster than the eager pipeline")
    ap.add_argument("--gather", default="auto", choices=["auto", "push", "p2p", "nccl"], help="tile exchange (N > 1)")
    ap.add_argument("--push-ctas", type=int, default=0, help="CTAs of the push kernel (0 = library default)")
    ap.add_argument("--ref-budget", type=float, default=300.0, help="--impl reference: seconds for all steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)
        # The JSON line is out and every stream is synchronised.  Leave without interpreter teardown: destructors of
        # CUDA-graph memory pools and of peer-mapped (symmetric) memory have been seen to abort there after a failed
        # capture, which would turn a finished measurement into a non-zero exit code.
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
