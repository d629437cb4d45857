"""GPU parity tests for hot path B (aperture-field assembly) through the C-ABI.

Oracle = outputs of the unmodified reference's build_nearfield on the synthetic lens
(tests/golden/nearfield_*.npz) and oracle/nearfield_oracle.py for sizes without a fixture.
The kernel computes in float64 and the drop-in returns complex128, so the tolerance here is far
below north_star's 1e-5; the complex64 device output is checked at 1e-6.
"""
import os

import numpy as np
import pytest

import synth_lens
from parity import field_error
from test_oracle_nearfield import CASES, library, periphery_from

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
inf = float("inf")


def call(fn, name, golden_dir, **extra):
    spec, kw = CASES[name]
    g = np.load(os.path.join(golden_dir, "nearfield_%s.npz" % name))
    collections, hgs = library(spec)
    sx, sy, sz = g["source"]
    explicit = name.endswith("ragged")
    kw = dict(kw, **extra)
    res = fn(source_x=sx, source_y=sy, source_z=sz, source_pol=str(g["pol"]), wavelength=float(g["wavelength"]),
             lens_periphery_summary=periphery_from(g, collections), lens_center_summary=g["center"],
             hexgridset=hgs, x_pts=g["x_pts"] if explicit else None, y_pts=g["y_pts"] if explicit else None, **kw)
    return res, g


@pytest.mark.parametrize("name", sorted(CASES))
def test_build_nearfield_matches_reference(name, golden_dir):
    from metalens_b200.nearfield import build_nearfield
    res, g = call(build_nearfield, name, golden_dir)
    escale = max(np.abs(g["Ex"]).max(), np.abs(g["Ey"]).max())
    hscale = max(np.abs(g["Hx"]).max(), np.abs(g["Hy"]).max())
    for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
        assert res[k].dtype == np.complex128 and res[k].shape == g[key].shape
        err = np.abs(res[k] - g[key]).max() / (escale if k < 2 else hscale)
        assert err < 1e-9, (key, err)
    np.testing.assert_array_equal(res[4], g["x_pts"])
    np.testing.assert_array_equal(res[5], g["y_pts"])
    assert abs(res[6] - g["power"]) <= 1e-11 * abs(g["power"])
    assert res[7] == g["n_glass"]


def test_complex64_device_output_and_big(golden_dir):
    """The device-resident complex64 output (what feeds NF->FF) and build_nearfield_big."""
    from metalens_b200.nearfield import NearfieldPlan, build_nearfield_big
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_ragged.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    periph = periphery_from(g, collections)
    plan = NearfieldPlan(580e-9, periph, g["center"], hgs)
    out, power = plan.run(0.0, 0.0, float(g["source"][2]), "x", g["x_pts"], g["y_pts"])
    assert out.dtype == torch.complex64
    got = out[:, :, :g["y_pts"].size].cpu().numpy()
    for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
        assert field_error(got[k], g[key]) < 3e-6      # fp32 order terms, complex64 output
    assert abs(power.item() - g["power"]) <= 1e-11 * abs(g["power"])
    res = build_nearfield_big(0.0, 0.0, float(g["source"][2]), "x", 580e-9, periph, g["center"], hgs,
                              x_pts=g["x_pts"], y_pts=g["y_pts"])
    for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
        assert field_error(res[k], g[key]) < 1e-9


def test_bounds_violation_raises_reference_valueerror(golden_dir):
    from metalens_b200.nearfield import build_nearfield
    e = np.load(os.path.join(golden_dir, "nearfield_error_small_plane.npz"))
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_onaxis.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    with pytest.raises(ValueError) as ei:
        build_nearfield(0.0, 0.0, -inf, "x", 580e-9, periphery_from(g, collections), g["center"], hgs)
    assert ei.value.args[0] == str(e["message"])
    assert ei.value.args[1] == float(e["value"]) and ei.value.args[2] == float(e["bound"])


def test_argument_assertions(golden_dir):
    from metalens_b200 import grating, lens_center
    from metalens_b200.nearfield import build_nearfield
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_onaxis.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    periph = periphery_from(g, collections)
    with pytest.raises(AssertionError):
        build_nearfield(0, 0, +1e-6, "x", 580e-9, periph, g["center"], hgs)          # source_z >= 0
    with pytest.raises(AssertionError):
        build_nearfield(0, 0, -1e-5, "q", 580e-9, periph, g["center"], hgs)          # bad polarisation
    with pytest.raises(AssertionError):
        build_nearfield(0, 0, -1e-5, "x", 580e-9, periph, g["center"], hgs,
                        x_pts=np.linspace(-1e-5, 1e-5, 20), y_pts=np.linspace(-1e-5, 1e-5, 20))   # dx > lambda/2
    # n_glass == 0 -> grating.n_glass(532) is not tabulated -> ValueError('bad wavelength...') (Q6)
    c0, h0 = synth_lens.make_library(grating, lens_center, synth_lens.SMALL_LENS, n_glass=0)
    with pytest.raises(ValueError):
        build_nearfield(0, 0, -1.43e-5, "x", 532e-9, periphery_from(g, c0), g["center"], h0)


def test_larger_lens_against_oracle():
    """A lens without a fixture (wider aperture, off-axis z-polarised source): CUDA vs numpy oracle,
    and the empty-grid early return."""
    from oracle import nearfield_oracle as no
    from metalens_b200 import grating, lens_center
    from metalens_b200.design import make_design
    from metalens_b200.nearfield import build_nearfield
    spec = dict(bands=[(15.0, 25.0, 1000e-9, 0.3), (25.0, 45.0, 650e-9, 1.1)], source_distance=30e-6, radius=29e-6)
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph, center, r_switch = make_design(collections, spec["source_distance"], spec["radius"], hgs)
    args = (1.1e-6, -0.6e-6, -30e-6, "z", 580e-9, periph, center, hgs)
    got = build_nearfield(*args)
    ref = no.build_nearfield(*args)
    assert got[0].shape == ref[0].shape and got[0].shape[0] >= 216
    # The default grid has an odd count here, so some samples sit exactly on symmetry lines:
    #  (1) centre samples exactly equidistant from two hex cells (y = 0 between cells at +-pitch/2): the reference takes
    #      whichever cell scipy's cKDTree meets first; build_nearfield resolves exactly those samples (the kernel reports
    #      them) with the same cKDTree call, so they must MATCH, not be skipped;
    #  (2) periphery samples exactly on the boundary between two grating copies, phi/angle_per_grating = k + 1/2:
    #      counted below and compared like every other sample.
    X, Y = np.meshgrid(got[4], got[5], indexing="ij")
    r = np.hypot(X, Y)
    d2 = (X.ravel()[:, None] - center[None, :, 0]) ** 2 + (Y.ravel()[:, None] - center[None, :, 1]) ** 2
    in_center = r.ravel() <= periph["r_min_list"][0]
    tied = (((d2 <= d2.min(axis=1, keepdims=True)).sum(axis=1) > 1) & in_center).reshape(X.shape)
    ring = np.searchsorted(np.hstack((periph["r_min_list"], periph["r_max_list"][-1])), r) - 1
    ring[ring == len(periph["r_min_list"])] = -1
    turns = np.arctan2(Y, X) / (2 * np.pi / periph["num_around_circle_list"][np.maximum(ring, 0)])
    on_wedge_edge = (np.abs(np.abs(turns - np.round(turns)) - 0.5) < 1e-9) & (ring >= 0)
    assert tied.sum() > 0 and on_wedge_edge.sum() > 0
    for k in range(4):
        assert field_error(got[k], ref[k]) < 1e-9                      # every sample, ties included
        assert field_error(got[k] * tied, ref[k] * tied) < 1e-9
    # without the tie resolution the same samples DIFFER (the rule "highest row" is not cKDTree's): the resolution is doing
    # something, and only on those samples
    from metalens_b200.nearfield import NearfieldPlan
    import torch as _t
    plan = NearfieldPlan(580e-9, periph, center, hgs)
    fast, _p = plan.run(1.1e-6, -0.6e-6, -30e-6, "z", got[4], got[5], out_dtype=_t.complex128, ties="fast")
    fast = fast[:, :, :got[5].size].cpu().numpy()
    differs = np.zeros(X.shape, bool)
    for k in range(4):
        differs |= np.abs(fast[k] - ref[k]) > 1e-9 * np.abs(ref[k]).max()
    assert not (differs & ~(tied | on_wedge_edge)).any()
    exact, _p = plan.run(1.1e-6, -0.6e-6, -30e-6, "z", got[4], got[5], out_dtype=_t.complex128, ties="reference")
    assert plan.last_tie_classes[0] == int(tied.sum()) and plan.last_tie_classes[1] >= int(on_wedge_edge.sum())
    assert abs(got[6] - ref[6]) <= 1e-11 * abs(ref[6])
    far = np.linspace(40e-6, 44e-6, 20)
    z = build_nearfield(*args, x_pts=far, y_pts=far)
    assert all(np.all(z[k] == 0) for k in range(4)) and z[6] == 0


@pytest.mark.parametrize("name", ["mid_z_offaxis", "mid_y_onaxis"])
def test_mid_size_lens_matches_reference(name, golden_dir):
    """The survey's probe-sized lens (720^2 / 675^2 samples, 17 rings, 1.26e5 hex cells) against values of the UNMODIFIED
    reference (tests/golden/make_nearfield_mid_golden.py): 24000 random samples plus EVERY exact nearest-cell tie (148 on
    the odd on-axis grid) -- ties are asserted, not masked -- incident power and the field energies."""
    from metalens_b200 import grating, lens_center
    from metalens_b200.design import make_design
    from metalens_b200.nearfield import build_nearfield
    g = np.load(os.path.join(golden_dir, "nearfield_%s.npz" % name))
    spec = synth_lens.MID_LENS
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph, center, _ = make_design(collections, spec["source_distance"], spec["radius"], hgs)
    assert center.shape == g["center"].shape and np.array_equal(center[:, 2], g["center"][:, 2])
    np.testing.assert_allclose(center[:, :2], g["center"][:, :2], rtol=1e-14, atol=1e-22)
    sx, sy, sz = g["source"]
    explicit = g["shape"][0] != 720
    res = build_nearfield(sx, sy, sz, str(g["pol"]), float(g["wavelength"]), periphery_from(g, collections), g["center"],
                          hgs, x_pts=g["x_pts"] if explicit else None, y_pts=g["y_pts"] if explicit else None)
    assert res[0].shape == tuple(g["shape"])
    np.testing.assert_array_equal(res[4], g["x_pts"])
    idx = g["index"]
    n_tie = int(g["tie"].sum())
    assert n_tie == (148 if explicit else 0)
    for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
        scale = float(g["scale_E"] if k < 2 else g["scale_H"])
        got = res[k].ravel()[idx]
        assert np.abs(got - g[key]).max() / scale < 1e-9, key
        if n_tie:
            assert np.abs(got[g["tie"]] - g[key][g["tie"]]).max() / scale < 1e-9, (key, "ties")
        assert abs(np.sum(np.abs(res[k]) ** 2) - g["sum_abs2"][k]) <= 1e-9 * g["sum_abs2"][k]     # whole-array energy
    assert abs(res[6] - float(g["power"])) <= 1e-11 * abs(float(g["power"]))


def test_plan_cache_distinguishes_wavelengths_within_one_nm(golden_dir):
    """Two build_nearfield calls on the same objects with wavelengths that round to the same nm (the table key,
    nearfield.py:86) must each use their own exact wavelength for k_vac / k_glass."""
    from oracle import nearfield_oracle as no
    from metalens_b200.nearfield import build_nearfield
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_onaxis.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    periph = periphery_from(g, collections)
    f = float(g["source"][2])
    x = np.linspace(-12e-6, 12e-6, 96)
    for wl in (580.0e-9, 580.4e-9):
        got = build_nearfield(0.0, 0.0, f, "x", wl, periph, g["center"], hgs, x_pts=x, y_pts=x)
        ref = no.build_nearfield(0.0, 0.0, f, "x", wl, periph, g["center"], hgs, x_pts=x, y_pts=x)
        for k in range(4):
            assert field_error(got[k], ref[k]) < 1e-9


def test_table_callable_matches_oracle():
    """T4: an AmplitudeTable is callable like the reference's RegularGridInterpolator."""
    from oracle import nearfield_oracle as no
    collections, hgs = library(synth_lens.SMALL_LENS)
    gc = collections[1][1]
    key = sorted(gc.interpolators, key=repr)[3]
    f = gc.interpolators[key]
    rng = np.random.default_rng(0)
    lo = np.array([g[0] for g in f.grid]); hi = np.array([g[-1] for g in f.grid])
    pts = lo + (hi - lo) * rng.random((500, 3))
    pts[0] = lo; pts[1] = hi; pts[2] = [f.grid[0][2], f.grid[1][1], f.grid[2][3]]      # nodes and corners
    got = f(pts)
    ref = no.trilinear(f.grid, f.values, pts)
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()
    with pytest.raises(ValueError):
        f(np.array([[hi[0] * 1.01 + 1, lo[1], lo[2]]]))


def test_lens_to_far_field_pipeline_on_device(golden_dir):
    """SURVEY KAT-5 / BASELINE cfg4 shape: GratingCollection + HexGridSet lens -> aperture fields
    (complex64, stays on the GPU) -> far-field power map, against the reference chain restated by the
    oracles (build_nearfield -> fft2(fftshift) -> farfield_from_nearfield)."""
    from oracle import farfield_oracle as fo
    from oracle import nearfield_oracle as no
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.nearfield import NearfieldPlan
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_onaxis.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    periph = periphery_from(g, collections)
    f = float(-g["source"][2])
    # power-of-two grid over the lens so the FFT path is exercised; spacing < lambda/2
    x = np.linspace(-12e-6, 12e-6, 128)
    nf = NearfieldPlan(580e-9, periph, g["center"], hgs)
    fields, p_in = nf.run(0.0, 0.0, -f, "x", x, x)
    ff = FarfieldPlan((128, 128), x[1] - x[0], x[1] - x[0], 580e-9, nf.n_glass, stride=1)
    assert ff.method == "fft"
    P, total = ff.run([fields[i][:, :128] for i in range(4)])
    ref_fields = no.build_nearfield(0.0, 0.0, -f, "x", 580e-9, periph, g["center"], hgs, x_pts=x, y_pts=x)
    P_ref, total_ref, *_ = fo.farfield_reference_path(*ref_fields[:4], x, x, 580e-9, ref_fields[7])
    from parity import power_map_error
    assert power_map_error(P.cpu().numpy(), P_ref) < 1e-5
    assert abs(total.item() - total_ref) <= 1e-5 * abs(total_ref)
    assert abs(p_in.item() - ref_fields[6]) <= 1e-11 * abs(ref_fields[6])
    # the lens transmits most of the incident power into propagating far-field bins
    assert 0.05 < total.item() / p_in.item() < 1.5


def test_incoherent_xyz_dipoles_on_device(golden_dir):
    """isotropic source = incoherent sum of x-, y- and z-polarised dipoles (nearfield.py:69-73):
    the accumulated device far field equals the sum of the three separate ones."""
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.nearfield import NearfieldPlan
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_onaxis.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    periph = periphery_from(g, collections)
    f = float(-g["source"][2])
    x = np.linspace(-12e-6, 12e-6, 128)
    nf = NearfieldPlan(580e-9, periph, g["center"], hgs)
    ff = FarfieldPlan((128, 128), x[1] - x[0], x[1] - x[0], 580e-9, nf.n_glass, stride=1)
    sets, singles, totals = [], [], 0.0
    for pol in "xyz":
        fields, _ = nf.run(0.0, 0.0, -f, pol, x, x)
        fs = [fields[i][:, :128].clone() for i in range(4)]
        sets.append(fs)
        P, t = ff.run(fs)
        singles.append(P.clone())
        totals += t.item()
    P_sum, total = ff.run_incoherent(sets)
    ref = singles[0] + singles[1] + singles[2]
    fin = torch.isfinite(ref)
    assert bool((torch.isnan(P_sum) == torch.isnan(ref)).all())
    assert ((P_sum - ref).abs()[fin].max() / ref[fin].max()).item() < 1e-6
    assert abs(total.item() - totals) <= 1e-12 * abs(totals)


def test_ring_records_and_warp_tiles(golden_dir):
    """mlb_nearfield_prepare: the per-ring records and the ring bin table equal their numpy definitions
    (nearfield.py:125, :161-165; scipy RGI find_indices for the period axis); and the assembly kernel gives the
    same fields for every warp tile shape (32x1 ... 4x8 samples) and register-budget variant."""
    from metalens_b200 import _lib
    from metalens_b200.nearfield import NearfieldPlan
    lib = _lib.load()
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_ragged.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    periph = periphery_from(g, collections)
    plan = NearfieldPlan(580e-9, periph, g["center"], hgs)
    torch.cuda.synchronize()
    n = plan.n_rings
    aux = plan._keep["ring_aux"].cpu().numpy().reshape(n, 8)
    ints = plan._keep["ring_aux"].cpu().view(torch.int32).numpy().reshape(n, 16)[:, 14:16]      # (gc, i2)
    rc = np.asarray(periph["r_center_list"], float); gp = np.asarray(periph["grating_period_list"], float)
    apg = 2.0 * np.pi / np.asarray(periph["num_around_circle_list"], float)
    assert np.array_equal(aux[:, 0], rc) and np.array_equal(aux[:, 1], gp) and np.array_equal(aux[:, 2], apg)
    assert np.array_equal(aux[:, 3], rc * apg) and np.array_equal(aux[:, 4], 2.0 * np.pi / gp)
    assert np.array_equal(aux[:, 5], 2.0 * np.pi / (rc * apg))
    gci = np.asarray(periph["gratingcollection_index_here_list"])
    assert np.array_equal(ints[:, 0], gci)
    for r in range(n):
        axis = plan.packs[gci[r]].axes[-plan.packs[gci[r]].n[2]:]
        i2 = int(np.clip(np.searchsorted(axis, gp[r], side="right") - 1, 0, axis.size - 2))
        assert ints[r, 1] == i2 and aux[r, 6] == (gp[r] - axis[i2]) / (axis[i2 + 1] - axis[i2])
    bounds = np.hstack((periph["r_min_list"], periph["r_max_list"][-1]))
    lut = plan._keep["ring_lut"].cpu().numpy()
    edges = np.arange(plan.n_lut) * (plan.lens_max_r / plan.n_lut)
    assert np.array_equal(lut[:-1], np.searchsorted(bounds, edges, side="left")) and lut[-1] == n + 1
    ref = None
    try:
        for tune in (105, 104, 103, 102):
            _lib.check(lib.mlb_nearfield_tune(tune), "tile")
            for variant in (6, 1):
                _lib.check(lib.mlb_nearfield_tune(variant), "variant")
                out, power = plan.run(0.0, 0.0, float(g["source"][2]), "x", g["x_pts"], g["y_pts"])
                got = out[:, :, :g["y_pts"].size].cpu().numpy()
                for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
                    assert field_error(got[k], g[key]) < 3e-6
                assert abs(power.item() - g["power"]) <= 1e-11 * abs(g["power"])
                if ref is None:
                    ref = got
                assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
            out64, _ = plan.run(0.0, 0.0, float(g["source"][2]), "x", g["x_pts"], g["y_pts"], out_dtype=torch.complex128)
            got64 = out64[:, :, :g["y_pts"].size].cpu().numpy()
            for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
                assert field_error(got64[k], g[key]) < 1e-9
    finally:
        lib.mlb_nearfield_tune(103)
        lib.mlb_nearfield_tune(6)
