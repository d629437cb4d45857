"""CPU: pin oracle/nearfield_oracle.py against the UNMODIFIED reference's build_nearfield
outputs (tests/golden/nearfield_*.npz from tests/golden/make_nearfield_golden.py), using
metalens_b200's own Grating/GratingCollection/HexGridSet table builders (T1-T3)."""
import os

import numpy as np
import pytest

import synth_lens
from metalens_b200 import grating, lens_center
from oracle import nearfield_oracle as no
from parity import field_error

inf = float("inf")


def periphery_from(g, collections):
    keys = ["r_center_list", "r_min_list", "r_max_list", "grating_period_list",
            "gratingcollection_index_here_list", "num_around_circle_list"]
    d = {k: g["periph_" + k] for k in keys}
    d["gratingcollection_list"] = [c[1] for c in collections]
    return d


CASES = {
    "small_x_onaxis": (synth_lens.SMALL_LENS, {}),
    "small_y_offaxis": (synth_lens.SMALL_LENS, {}),
    "small_z_onaxis": (synth_lens.SMALL_LENS, {}),
    "plane_x": (synth_lens.PLANE_LENS, {"dipole_moment": 1.0}),
    "plane_lens_y_point": (synth_lens.PLANE_LENS, {}),
    "small_x_ragged": (synth_lens.SMALL_LENS, {}),
}
_LIB = {}


def library(spec):
    key = id(spec)
    if key not in _LIB:
        _LIB[key] = synth_lens.make_library(grating, lens_center, spec)
    return _LIB[key]


def run_case(fn, name, golden_dir):
    spec, kw = CASES[name]
    g = np.load(os.path.join(golden_dir, "nearfield_%s.npz" % name))
    collections, hgs = library(spec)
    sx, sy, sz = g["source"]
    explicit = name.endswith("ragged")
    res = fn(source_x=sx, source_y=sy, source_z=sz, source_pol=str(g["pol"]), wavelength=float(g["wavelength"]),
             lens_periphery_summary=periphery_from(g, collections), lens_center_summary=g["center"],
             hexgridset=hgs, x_pts=g["x_pts"] if explicit else None, y_pts=g["y_pts"] if explicit else None, **kw)
    return res, g


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference(name, golden_dir):
    res, g = run_case(no.build_nearfield, name, golden_dir)
    scale = max(np.abs(g["Ex"]).max(), np.abs(g["Ey"]).max())
    hscale = max(np.abs(g["Hx"]).max(), np.abs(g["Hy"]).max())
    for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
        err = np.abs(res[k] - g[key]).max() / (scale if k < 2 else hscale)
        assert err < 1e-11, (key, err)
    np.testing.assert_array_equal(res[4], g["x_pts"])
    np.testing.assert_array_equal(res[5], g["y_pts"])
    assert abs(res[6] - g["power"]) <= 1e-12 * abs(g["power"])
    assert res[7] == g["n_glass"]


@pytest.mark.parametrize("name", ["mid_z_offaxis", "mid_y_onaxis"])
def test_oracle_matches_mid_reference(name, golden_dir):
    """The probe-sized lens of SURVEY section 8 (17 rings, 1.26e5 hex cells, off-axis z-dipole / odd on-axis grid with 148
    exact nearest-cell ties): the oracle against the stored subset of the UNMODIFIED reference's output."""
    g = np.load(os.path.join(golden_dir, "nearfield_%s.npz" % name))
    collections, hgs = synth_lens.make_library(grating, lens_center, synth_lens.MID_LENS)
    sx, sy, sz = g["source"]
    explicit = g["shape"][0] != 720
    res = no.build_nearfield(sx, sy, sz, str(g["pol"]), float(g["wavelength"]), periphery_from(g, collections),
                             g["center"], hgs, x_pts=g["x_pts"] if explicit else None,
                             y_pts=g["y_pts"] if explicit else None)
    idx = g["index"]
    for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
        scale = float(g["scale_E"] if k < 2 else g["scale_H"])
        assert np.abs(res[k].ravel()[idx] - g[key]).max() / scale < 1e-11, key
    assert abs(res[6] - float(g["power"])) <= 1e-12 * abs(float(g["power"]))


def test_oracle_big_equals_single(golden_dir):
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_ragged.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    args = dict(source_x=0.0, source_y=0.0, source_z=float(g["source"][2]), source_pol="x", wavelength=580e-9,
                lens_periphery_summary=periphery_from(g, collections), lens_center_summary=g["center"],
                hexgridset=hgs, x_pts=g["x_pts"], y_pts=g["y_pts"])
    a = no.build_nearfield(**args)
    b = no.build_nearfield_big(pts_at_a_time=105 * 20, **args)
    for k in range(4):
        assert field_error(b[k], a[k]) < 1e-14
    assert abs(a[6] - b[6]) < 1e-12 * abs(a[6])


def test_oracle_bounds_error_matches_reference(golden_dir):
    """Normal incidence is outside the SMALL_LENS tables: same ValueError args as the reference."""
    e = np.load(os.path.join(golden_dir, "nearfield_error_small_plane.npz"))
    g = np.load(os.path.join(golden_dir, "nearfield_small_x_onaxis.npz"))
    collections, hgs = library(synth_lens.SMALL_LENS)
    with pytest.raises(ValueError) as ei:
        no.build_nearfield(0.0, 0.0, -inf, "x", 580e-9, periphery_from(g, collections), g["center"], hgs)
    assert ei.value.args[0] == str(e["message"])
    assert ei.value.args[1] == float(e["value"]) and ei.value.args[2] == float(e["bound"])


def test_good_fft_number():
    assert [no.good_fft_number(n) for n in (1, 7, 91, 675, 676, 4097)] == [1, 8, 96, 675, 720, 4320]


def test_table_builders_shapes():
    collections, hgs = library(synth_lens.SMALL_LENS)
    gc = collections[0][1]
    key = next(iter(gc.interpolators))
    f = gc.interpolators[key]
    assert f.values.shape == (5, 5, 7) and len(gc.interpolator_bounds) == 6
    assert np.array_equal(f.values[:, :, 0], f.values[:, :, 1]) and np.array_equal(f.values[:, :, -1], f.values[:, :, -2])
    assert gc.interpolator_bounds[4] == 0.99 * gc.grating_list[0].grating_period
    hk = next(iter(hgs.interpolators))
    assert hgs.interpolators[hk].values.shape == (5, 5, 20)
    assert hgs.interpolator_bounds[4:] == (0, 19)
    with pytest.raises(ValueError):
        lens_center.HexGridSet(sep=320e-9, cyl_height=550e-9, grating_list=hgs.grating_list).build_interpolators()
    with pytest.raises(ValueError):
        grating.n_glass(532)
    assert grating.n_glass(580) == 1.459
    g2 = gc.grating_list[0].copy()
    assert g2.grating_period == pytest.approx(gc.grating_list[0].grating_period, rel=1e-12) and len(g2.data) == len(gc.grating_list[0].data)


def test_packed_library_round_trip(tmp_path, golden_dir):
    """collections saved in the packed .npz format come back with identical tables, and the
    oracle near field computed from the re-loaded library equals the reference fixture."""
    from metalens_b200 import tables
    collections, hgs = library(synth_lens.SMALL_LENS)
    loaded = []
    for i, (band, gc) in enumerate(collections):
        path = str(tmp_path / ("gc%d.npz" % i))
        tables.save_library(path, gc)
        back = tables.load_library(path)
        assert sorted(back.interpolators, key=repr) == sorted(gc.interpolators, key=repr)
        for k in gc.interpolators:
            assert np.array_equal(back.interpolators[k].values, gc.interpolators[k].values)
            assert all(np.array_equal(a, b) for a, b in zip(back.interpolators[k].grid, gc.interpolators[k].grid))
        assert back.interpolator_bounds == tuple(gc.interpolator_bounds)
        assert back.lateral_period == gc.lateral_period and back.lens_type == "round"
        loaded.append([band, back])
    path = str(tmp_path / "hgs.npz")
    tables.save_library(path, hgs)
    hgs2 = tables.load_library(path)
    assert np.array_equal(hgs2.x_amp_list, hgs.x_amp_list) and hgs2.interpolator_bounds == tuple(hgs.interpolator_bounds)
    g = np.load(os.path.join(golden_dir, "nearfield_small_y_offaxis.npz"))
    sx, sy, sz = g["source"]
    res = no.build_nearfield(sx, sy, sz, "y", 580e-9, periphery_from(g, loaded), g["center"], hgs2)
    scale = max(np.abs(g["Ex"]).max(), np.abs(g["Ey"]).max())
    assert np.abs(res[0] - g["Ex"]).max() / scale < 1e-11 and np.abs(res[1] - g["Ey"]).max() / scale < 1e-11
