"""BASELINE config 4 shape on one GPU: NA~0.94 synthetic lens (GratingCollection x3 + HexGridSet) on an
8192^2 aperture built by the fused assembly kernel, then NF->FF on (a) every 4th bin (2048^2) and (b) all
8192^2 bins with the FFT passes.  Prints timings and a cross-check of the FFT path against the dense
tiled reduction on a handful of bins.  Not a bench line (bench.py is); evidence for DESIGN.md."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth_lens
from metalens_b200 import grating, lens_center
from metalens_b200.design import make_design
from metalens_b200.nearfield import NearfieldPlan
from metalens_b200.farfield import FarfieldPlan

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = 580e-9
R = M * (wl / 2.2) / 2
f = R / math.tan(math.asin(0.94))
spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1), (45.0, 71.0, 300e-9, 2.3)],
            source_distance=f, radius=R * 0.999)
t0 = time.time()
collections, hgs = synth_lens.make_library(grating, lens_center, spec)
periph, center, _ = make_design(collections, f, spec["radius"], hgs)
print("design: %d rings, %d hex cells, %.1f s host" % (len(periph["r_min_list"]), len(center), time.time() - t0), flush=True)
nf = NearfieldPlan(wl, periph, center, hgs)
x = np.linspace(-R, R, M)
out = torch.zeros((4, M, M), dtype=torch.complex64, device="cuda")

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_nf = timed(lambda: nf.run(0.0, 0.0, -f, "x", x, x, out=out))
print("assembly %dx%d: %.2f ms (%.2e samples/s, %.0f GB/s written)" % (M, M, t_nf, M * M / t_nf * 1e3, 32 * M * M / t_nf / 1e6), flush=True)
fields = [out[i] for i in range(4)]
d = float(x[1] - x[0])
for stride in (4, 1):
    plan = FarfieldPlan((M, M), d, d, wl, nf.n_glass, stride=stride)
    t = timed(lambda: plan.run(fields))
    K = plan.Kx
    P, total = plan.run(fields)
    P = P.clone()          # the per-kernel timings below reuse (and for N >= 4096 scratch over) the plan's buffers
    print("NF->FF stride %d (%s): %dx%d far field in %.3f ms = %.2e points/s; total_P/P_in = %.4f" % (
        stride, plan.method, K, K, t, K * K / t * 1e3, total.item() / nf.run(0.0, 0.0, -f, "x", x, x, out=out)[1].item()), flush=True)
    for name, fn, nbytes, flops in plan.steps(fields):
        tk = timed(fn)
        print("    %-14s %.3f ms  %.0f GB/s" % (name, tk, nbytes / tk / 1e6), flush=True)
    # cross-check a few bins against the dense tiled reduction (independent kernels)
    ii = np.array([K // 2, K // 2 + 3, K // 3, K // 2 + K // 7]); jj = np.array([K // 2, K // 2 - 5, K // 2 + K // 5])
    dense = FarfieldPlan((M, M), d, d, wl, nf.n_glass, ux=plan.ux[ii], uy=plan.uy[jj], method="dense")
    Pd = dense.run(fields)[0].cpu().numpy()
    Pf = P.cpu().numpy()[np.ix_(ii, jj)]
    print("    fft vs dense on %d bins: max|dP|/max|P| = %.2e" % (Pd.size, np.nanmax(np.abs(Pd - Pf)) / np.nanmax(P.cpu().numpy())), flush=True)
    del plan, dense, P
    torch.cuda.empty_cache()
