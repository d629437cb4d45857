"""GPU: BASELINE configurations 2, 3 and 4 at FULL size against the oracle -- complete maps, every batch item.

cfg2 / cfg3: the oracle runs the reference path (4 x fft2(fftshift), radiated-power helper; all M^2 bins, threaded but
bit-identical to the single-thread path) and is sub-sampled at the stride; every one of the K x K bins of every
wavelength / polarisation is compared, NaN masks included.
cfg4: the synthetic NA 0.94 lens.  At 2048^2 the whole chain (assembly -> far field, strided and all bins) is compared
with the oracle chain end to end; at 8192^2 the assembly is compared on 48 random aperture rows, the far field on
a random 100 x 100 lattice of bins (1e4 bins, float64 separable DFT of the full aperture), and the one-aperture-over-
8-ranks path (virtual ranks) must reproduce the single-GPU map bit for bit.
"""
import math
import os

import numpy as np
import pytest

import apertures
from parity import FF_TOL, field_error, power_map_error

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CONFIGS = {
    "cfg2": dict(M=2048, stride=4, items=[(532e-9, 1.4607, False), (532e-9, 1.4607, True)]),
    "cfg3": dict(M=4096, stride=4, items=[(450e-9, 1.466, False), (532e-9, 1.4607, False), (635e-9, 1.457, False)]),
}


def _workers():
    try:
        return max(1, min(32, len(os.sched_getaffinity(0))))
    except AttributeError:
        return 4


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_full_map_of_every_item(name):
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import FarfieldPlan
    cfg = CONFIGS[name]
    M, s = cfg["M"], cfg["stride"]
    K = M // s
    for i, (wl, ng, rot) in enumerate(cfg["items"]):
        Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 1000 + i, wl, ng, rotate=rot)     # the bench's apertures
        d = float(x[1] - x[0])
        plan = FarfieldPlan((M, M), d, d, wl, ng, stride=s)
        assert plan.method == "fft"
        P, total = plan.run([torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)])
        P = P.cpu().numpy()
        P_ref, total_ref, ux, uy, dux, duy = fo.farfield_reference_path_threads(Ex, Ey, Hx, Hy, x, y, wl, ng,
                                                                                workers=_workers())
        sub = P_ref[::s, ::s]
        assert P.shape == (K, K) == sub.shape
        np.testing.assert_array_equal(plan.ux, ux.ravel()[::s])
        np.testing.assert_array_equal(plan.uy, uy.ravel()[::s])
        assert power_map_error(P, sub) < FF_TOL, (name, i)                      # all K^2 bins, identical NaN mask
        total_sub = sub[np.isfinite(sub)].sum() * (s * dux) * (s * duy)
        assert abs(total.item() - total_sub) <= FF_TOL * abs(total_sub)
        del plan, P_ref


def test_strided_total_is_a_subsampled_sum():
    """total_P on an every-s-th-bin grid is the Riemann sum of P over those bins with the coarser cell s*du (what
    nearfield_farfield.py:74 does on its own grid).  It reproduces the stride-1 total when the far field is smooth on the
    scale of s bins (random aperture, a converging lens wave: ~1).  A COLLIMATED aperture -- what the design_collimator
    lens of cfg4 produces -- radiates into a diffraction-limited spot about one FFT bin wide that sits on a sampled bin:
    the coarse quadrature weights it with a cell s^2 times too large (the total_P / P_in = 1.37 of the 8192^2 lens at
    stride 4 against 0.91 at stride 1 in profiles/).  P itself is exact at every sampled bin in all cases."""
    from metalens_b200.farfield import FarfieldPlan
    M, s = 1024, 4
    wl, ng = 532e-9, apertures.N_GLASS[532]
    ratios = {}
    for kind in ("random", "converging", "collimated"):
        if kind == "random":
            Ex, Ey, Hx, Hy, x, y = apertures.gaussian_random(M, 3, wl)
        elif kind == "converging":
            Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 3, wl, ng, na=0.9, noise=0.0)
        else:
            Ex, Ey, Hx, Hy, x, y = apertures.disc(M, wl, ng, radius_samples=400)
            Ex, Ey, Hx, Hy = (a.astype(np.complex64) for a in (Ex, Ey, Hx, Hy))
        d = float(x[1] - x[0])
        dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (Ex, Ey, Hx, Hy)]
        full = FarfieldPlan((M, M), d, d, wl, ng, stride=1)
        P1, t1 = full.run(dev)
        P1 = P1.cpu().numpy().astype(np.float64)
        sub = FarfieldPlan((M, M), d, d, wl, ng, stride=s)
        Ps, ts = sub.run(dev)
        assert power_map_error(Ps.cpu().numpy(), P1[::s, ::s]) < FF_TOL
        coarse = P1[::s, ::s]
        riemann = coarse[np.isfinite(coarse)].sum() * sub.dux * sub.duy
        assert abs(ts.item() - riemann) <= 2e-5 * abs(riemann)
        ratios[kind] = ts.item() / t1.item()
    assert abs(ratios["random"] - 1.0) < 0.05 and abs(ratios["converging"] - 1.0) < 0.05
    assert ratios["collimated"] > 1.5


# ----------------------------------------------------------------------------- cfg4
def _cfg4_lens(M, wl=580e-9):
    import synth_lens
    from metalens_b200 import grating, lens_center
    from metalens_b200.design import make_design
    R = M * (wl / 2.2) / 2
    f = R / math.tan(math.asin(0.94))
    spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1), (45.0, 71.0, 300e-9, 2.3)],
                source_distance=f, radius=R * 0.999)
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph, center, _ = make_design(collections, f, spec["radius"], hgs)
    return periph, center, hgs, R, f


def test_cfg4_chain_full_maps_at_2048():
    """NA 0.94 lens on a 2048^2 grid: assembled fields (complex64) against the numpy oracle on every sample; far field
    on every 4th bin and on all bins against the oracle chain (oracle near field -> reference far-field path)."""
    from oracle import farfield_oracle as fo
    from oracle import nearfield_oracle as no
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.nearfield import NearfieldPlan
    M, wl = 2048, 580e-9
    periph, center, hgs, R, f = _cfg4_lens(M, wl)
    x = np.linspace(-R, R, M)
    nf = NearfieldPlan(wl, periph, center, hgs)
    fields, p_in = nf.run(0.3e-6, -0.2e-6, -f, "x", x, x, ties="reference")
    ref = no.build_nearfield_big(0.3e-6, -0.2e-6, -f, "x", wl, periph, center, hgs, x_pts=x, y_pts=x)
    got = fields[:, :, :M].cpu().numpy()
    for k in range(4):
        assert field_error(got[k], ref[k]) < 3e-6, k                  # complex64 output
    assert abs(p_in.item() - ref[6]) <= 1e-11 * abs(ref[6])
    P_ref, total_ref, *_ = fo.farfield_reference_path_threads(*ref[:4], x, x, wl, ref[7], workers=_workers())
    d = float(x[1] - x[0])
    dev = [fields[i][:, :M] for i in range(4)]
    for s in (4, 1):
        plan = FarfieldPlan((M, M), d, d, wl, nf.n_glass, stride=s)
        assert plan.method == "fft"
        P, total = plan.run(dev)
        assert power_map_error(P.cpu().numpy(), P_ref[::s, ::s]) < FF_TOL, s
    assert abs(total.item() - total_ref) <= FF_TOL * abs(total_ref)    # stride 1: the reference's own total_P


def test_cfg4_full_size_8192():
    from oracle import farfield_oracle as fo
    from oracle import nearfield_oracle as no
    from metalens_b200 import _lib
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.nearfield import NearfieldPlan
    from metalens_b200.peer import VirtualPeers
    from metalens_b200.slab import SlabFarfield, assemble_slab
    M, wl, s = 8192, 580e-9, 4
    periph, center, hgs, R, f = _cfg4_lens(M, wl)
    x = np.linspace(-R, R, M)
    d = float(x[1] - x[0])
    nf = NearfieldPlan(wl, periph, center, hgs)
    full = torch.zeros((4, M, M), dtype=torch.complex64, device="cuda")
    _, p_in = nf.run(0.0, 0.0, -f, "x", x, x, out=full)
    # hot path B on 6 random blocks of 8 aperture rows (all 8192 columns each: 3.9e5 samples) against the oracle
    rng = np.random.default_rng(4)
    for r0 in np.sort(rng.choice(M - 8, size=6, replace=False)):
        ref = no.build_nearfield(0.0, 0.0, -f, "x", wl, periph, center, hgs, x_pts=x[r0:r0 + 8], y_pts=x)
        blk, _p = nf.run(0.0, 0.0, -f, "x", x[r0:r0 + 8], x, ties="reference")      # index ties as the reference takes them
        got = blk.cpu().numpy()
        same = (blk.view(torch.float32) == full[:, r0:r0 + 8, :].contiguous().view(torch.float32)).all(dim=-1)
        assert float(same.float().mean()) > 0.999                                  # ... which touches a handful of samples
        scale_e = max(np.abs(ref[0]).max(), np.abs(ref[1]).max(), 1e-300)
        scale_h = max(np.abs(ref[2]).max(), np.abs(ref[3]).max(), 1e-300)
        for k in range(4):
            assert np.abs(got[k] - ref[k]).max() / (scale_e if k < 2 else scale_h) < 3e-6, (r0, k)
    # hot path A: single-GPU far field (same kernels as the ranks) ...
    lib = _lib.load()
    plan = FarfieldPlan((M, M), d, d, wl, nf.n_glass, stride=s, method="fft", fuse_power="always")
    old = lib.mlb_get_option(b"rows_engine")
    lib.mlb_set_option(b"rows_engine", 0)
    P1, total1 = plan.run([full[i] for i in range(4)])
    lib.mlb_set_option(b"rows_engine", old)
    P1, total1 = P1.clone(), total1.clone()
    # ... against the float64 separable DFT of the SAME aperture on a random 100 x 100 lattice of bins (1e4 bins)
    K = M // s
    ii = np.sort(rng.choice(K, size=100, replace=False))
    jj = np.sort(rng.choice(K, size=100, replace=False))
    host = full.cpu().numpy()
    P_ref, _ = fo.farfield_dense(host[0], host[1], host[2], host[3], d, d, plan.ux[ii], plan.uy[jj], wl, nf.n_glass)
    del host
    sub = P1.cpu().numpy()[np.ix_(ii, jj)]
    assert np.array_equal(np.isnan(sub), np.isnan(P_ref))
    fin = np.isfinite(P_ref)
    assert fin.sum() > 5000
    assert np.abs(sub - P_ref)[fin].max() / float(np.nanmax(P1.cpu().numpy())) < FF_TOL
    # ... and the same lens over 8 (virtual) ranks: rows of the assembly per rank, scattered row pass, column slabs
    world = 8
    vp = VirtualPeers(world)
    slabs = [SlabFarfield((M, M), d, d, wl, nf.n_glass, s, vp.view(r)) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    local, powers = [], []
    for r, sl in enumerate(slabs):
        loc, p = assemble_slab(nf, sl, (0.0, 0.0, -f), "x", x, x)
        idx = torch.from_numpy(sl.x_rows).cuda()
        assert torch.equal(loc.view(torch.float32), full.index_select(1, idx).contiguous().view(torch.float32))
        local.append(loc)
        powers.append(float(p))
    assert abs(sum(powers) - float(p_in)) <= 1e-12 * abs(float(p_in))
    del full
    for r, sl in enumerate(slabs):
        sl.warm([local[r][i] for i in range(4)])
    torch.cuda.synchronize()
    for r, sl in enumerate(slabs):
        with torch.cuda.stream(streams[r]):
            sl.run([local[r][i] for i in range(4)], wait=False)
    for r, sl in enumerate(slabs):
        with torch.cuda.stream(streams[r]):
            sl.finish()
    torch.cuda.synchronize()
    for sl in slabs:
        sl.chan.check()
        assert bool(((sl.P == P1) | (torch.isnan(sl.P) & torch.isnan(P1))).all())
        assert float(sl.total) == float(total1)
