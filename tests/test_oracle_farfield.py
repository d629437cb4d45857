"""CPU: pin oracle/farfield_oracle.py against outputs of the UNMODIFIED reference
(tests/golden/farfield_*.npz, made by tests/golden/make_farfield_golden.py)."""
import os

import numpy as np
import pytest

import apertures
from oracle import farfield_oracle as fo
from parity import power_map_error

WL = 532e-9
NG = apertures.N_GLASS[532]

CASES = {
    "kat1_uniform": lambda: apertures.uniform(128, WL, NG),
    "kat2_disc": lambda: apertures.disc(128, WL, NG),
    "kat3_tilted": lambda: apertures.tilted_te(128, WL, NG),
    "rand128_seed0": lambda: apertures.gaussian_random(128, 0, WL),
    "rand_48x40_seed5": lambda: apertures.gaussian_random(48, 5, WL, My=40),
    "rand_45x27_seed6": lambda: apertures.gaussian_random(45, 6, WL, My=27),
    "lens256_seed1": lambda: apertures.focusing_lens(256, 1, WL, NG),
    "lens256_seed1_rot": lambda: apertures.focusing_lens(256, 1, WL, NG, rotate=True),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_path_matches_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "farfield_%s.npz" % name))
    Ex, Ey, Hx, Hy, x, y = CASES[name]()
    P, total, ux, uy, dux, duy = fo.farfield_reference_path(Ex, Ey, Hx, Hy, x, y, WL, NG)
    assert power_map_error(P, g["P"]) < 1e-12
    assert abs(total - g["total_P"]) <= 1e-12 * abs(g["total_P"])
    np.testing.assert_allclose(ux.ravel(), g["ux"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(uy.ravel(), g["uy"], rtol=0, atol=1e-15)
    assert dux == g["dux"] and duy == g["duy"]


@pytest.mark.parametrize("name,rot", [("lens675_seed9", False), ("lens675_seed9_rot", True)])
def test_reference_path_matches_mid_golden(name, rot, golden_dir):
    """675 x 675: the reference's own typical default grid size (good_fft_number, odd, 3^3 5^2), sampled subset of the
    UNMODIFIED reference's map (tests/golden/make_farfield_mid_golden.py)."""
    g = np.load(os.path.join(golden_dir, "farfield_%s.npz" % name))
    wl, ng = float(g["wavelength"]), float(g["n_glass"])
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(675, 9, wl, ng, rotate=rot)
    P, total, ux, uy, dux, duy = fo.farfield_reference_path(Ex, Ey, Hx, Hy, x, y, wl, ng)
    got = P.ravel()[g["index"]]
    assert np.array_equal(np.isnan(got), np.isnan(g["P"])) and int(np.isnan(P).sum()) == int(g["nan_count"])
    fin = np.isfinite(g["P"])
    assert np.abs(got - g["P"])[fin].max() / float(g["P_max"]) < 1e-12
    assert abs(total - float(g["total_P"])) <= 1e-12 * abs(float(g["total_P"]))
    np.testing.assert_allclose(ux.ravel(), g["ux"], rtol=0, atol=1e-15)


def test_known_answers(golden_dir):
    """SURVEY KAT-1..3: energy conservation of the transform."""
    g = np.load(os.path.join(golden_dir, "farfield_kat1_uniform.npz"))
    assert abs(g["total_P"] / g["P_in"] - 1 / (1 + 1e-5)) < 1e-9      # uz+1e-5 regulariser at DC
    P = g["P"]
    assert np.nanargmax(P) == np.ravel_multi_index((64, 64), P.shape)
    g = np.load(os.path.join(golden_dir, "farfield_kat2_disc.npz"))
    assert abs(g["total_P"] / g["P_in"] - 1.000309) < 2e-6
    g = np.load(os.path.join(golden_dir, "farfield_kat3_tilted.npz"))
    assert abs(g["total_P"] / g["P_in"] - 0.99999) < 1e-5
    i, j = np.unravel_index(np.nanargmax(g["P"]), g["P"].shape)
    assert abs(g["ux"][i] - 0.3414) < 1e-3 and g["uy"][j] == 0


@pytest.mark.parametrize("name,stride", [("rand128_seed0", 4), ("rand_48x40_seed5", 1),
                                         ("rand_45x27_seed6", 1), ("lens256_seed1", 8)])
def test_dense_sum_matches_golden_bins(name, stride, golden_dir):
    """SURVEY KAT-4: the direct separable sum at (every stride-th) fftshifted bin
    equals the reference FFT result, including odd and non-square grids."""
    g = np.load(os.path.join(golden_dir, "farfield_%s.npz" % name))
    Ex, Ey, Hx, Hy, x, y = CASES[name]()
    ux, uy = g["ux"][::stride], g["uy"][::stride]
    P, _ = fo.farfield_dense(Ex, Ey, Hx, Hy, x[1] - x[0], y[1] - y[0], ux, uy, WL, NG)
    assert power_map_error(P, g["P"][::stride, ::stride]) < 1e-11


def test_grid_validation():
    x = np.arange(8) * 1e-7
    with pytest.raises(AssertionError):
        fo.check_uniform_axis(x * 10, WL)          # spacing >= lambda/2
    bad = x.copy(); bad[3] += 1e-9
    with pytest.raises(AssertionError):
        fo.check_uniform_axis(bad, WL)             # non-uniform


@pytest.mark.parametrize("workers", [1, 3, 8])
def test_threaded_reference_path_is_bit_identical(workers):
    """The all-cores flavour of the CPU baseline (bench.py) is the same arithmetic as the single-thread path:
    chunking the uy loop (nearfield_farfield.py:45-66) must not change a single bit, NaN mask included."""
    Ex, Ey, Hx, Hy, x, y = apertures.gaussian_random(96, 11, WL, My=70)
    a = fo.farfield_reference_path(Ex, Ey, Hx, Hy, x, y, WL, NG)
    b = fo.farfield_reference_path_threads(Ex, Ey, Hx, Hy, x, y, WL, NG, workers, points_at_a_time=96 * 13)
    assert np.array_equal(a[0], b[0], equal_nan=True) and a[1] == b[1]
    for u, v in zip(a[2:], b[2:]):
        assert np.array_equal(u, v)
