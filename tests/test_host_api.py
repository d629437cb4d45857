"""CPU: the reference-facing host surface behaves like the reference's own functions -- names, argument
meaning, return values, exception types and messages -- checked LIVE against the unmodified reference (installed copy,
baseline/_ref) where it is present, and against the reference's definitions restated here otherwise.  No kernel runs:
everything below happens before the first device call."""
import math

import numpy as np
import pytest

import synth_lens
from metalens_b200 import farfield, grating, lens_center, nearfield
from metalens_b200.units import nm, um

degree = math.pi / 180


@pytest.fixture(scope="module")
def ref():
    from baseline import install_ref
    mods = install_ref.load()
    if mods is None:
        pytest.skip("baseline/_ref not installed (run __graft_entry__.build() where /root/reference exists)")
    return mods


def _smooth_at_least(goal):
    """nearfield.py:30-36 by brute force: the smallest 2^a 3^b 5^c >= goal."""
    n = max(int(math.ceil(goal)), 1)
    while True:
        m = n
        for f in (2, 3, 5):
            while m % f == 0:
                m //= f
        if m == 1:
            return n
        n += 1


def test_good_fft_number_definition():
    for goal in list(range(1, 700)) + [1023, 1024, 1025, 3375, 3376, 4097, 8191, 8192, 8193, 19999]:
        assert nearfield.good_fft_number(goal) == _smooth_at_least(goal), goal
    # the default grid of build_nearfield (nearfield.py:95-97) passes a float goal
    assert nearfield.good_fft_number(674.3) == 675 and nearfield.good_fft_number(675.0) == 675


def test_good_fft_number_equals_reference(ref):
    theirs = ref["nearfield"].good_fft_number
    for goal in list(range(1, 2200)) + [3374.2, 4096, 4097, 8192.5, 12345]:
        assert nearfield.good_fft_number(goal) == theirs(goal), goal


def test_n_glass_table_and_error(ref):
    """grating.py:1274-1288: tabulated wavelengths only; same ValueError text."""
    for w in (450, 500, 525, 550, 575, 580, 600, 625, 650):
        assert grating.n_glass(w) == ref["grating"].n_glass(w)
    for w in (532, 635, 580.0 + 1e-9):
        with pytest.raises(ValueError) as ours:
            grating.n_glass(w)
        with pytest.raises(ValueError) as theirs:
            ref["grating"].n_glass(w)
        assert ours.value.args == theirs.value.args


def _some_grating(mod, with_data=True):
    g = mod.Grating(lateral_period=412.5 * nm, cyl_height=550 * nm, grating_period=1234.5 * nm, n_glass=1.459,
                    n_tio2=2.372, xyrra_list_in_nm_deg=np.array([[0., 10., 60., 80., 15.], [300., -40., 55., 55., 0.]]))
    if with_data:
        g.data = synth_lens.table_rows(580, [0.1, 0.3], [-0.2, 0.2], 1234.5 * nm, 412.5 * nm, salt=0.4)[:6]
    return g


def test_grating_repr_and_copy_round_trip(ref):
    """repr() is the reference's persistence format (README.md:29-34, grating.py:263-281): the same string, and
    eval(repr) -- copy() -- gives back the same object, in both directions."""
    ours, theirs = _some_grating(grating), _some_grating(ref["grating"])
    assert repr(ours) == repr(theirs)
    back = ours.copy()
    assert repr(back) == repr(ours) and back.data == ours.data
    assert np.array_equal(back.xyrra_list, ours.xyrra_list)
    # a string saved by the reference loads into this package's class (and vice versa)
    ns = {"Grating": grating.Grating, "nm": nm, "np": np, "array": np.array}
    loaded = eval(repr(theirs), ns)
    assert repr(loaded) == repr(theirs) and loaded.grating_period == theirs.grating_period
    bare_ours, bare_theirs = _some_grating(grating, False), _some_grating(ref["grating"], False)
    assert repr(bare_ours) == repr(bare_theirs)


def test_grating_constructor_rules(ref):
    """grating.py:121-127: either grating_period or (target_wavelength, angle_in_air)."""
    for mod in (grating, ref["grating"]):
        g = mod.Grating(lateral_period=400 * nm, cyl_height=550 * nm, target_wavelength=580 * nm,
                        angle_in_air=30 * degree)
        assert g.grating_period == 580 * nm / math.sin(30 * degree)
        assert g.get_angle_in_air(580 * nm) == math.asin(580 * nm / g.grating_period)
        with pytest.raises(AssertionError):
            mod.Grating(lateral_period=400 * nm, cyl_height=550 * nm, grating_period=1 * um, angle_in_air=0.3)
        with pytest.raises(ValueError):
            mod.Grating(lateral_period=400 * nm, cyl_height=550 * nm, grating_period=500 * nm).get_angle_in_air(580 * nm)


def test_collection_get_one_and_consistency(ref):
    """GratingCollection: sorted by period, round-lens consistency check (grating.py:955-969), get_one() with the
    pillar geometry interpolated between neighbours (grating.py:981-1047)."""
    built = {}
    for name, mod in (("ours", grating), ("theirs", ref["grating"])):
        gl = []
        for k, a in enumerate(np.linspace(20, 40, 5) * degree):
            gl.append(mod.Grating(lateral_period=900 * nm * math.tan(a), cyl_height=550 * nm,
                                  grating_period=580 * nm / math.sin(a), n_glass=1.459, n_tio2=2.372,
                                  xyrra_list_in_nm_deg=np.array([[10. * k, 5., 50. + k, 60. + 2 * k, 3. * k]])))
        gc = mod.GratingCollection(target_wavelength=580 * nm, lateral_period=900 * nm, lens_type="round",
                                   grating_list=gl[::-1])
        periods = [g.grating_period for g in gc.grating_list]
        assert periods == sorted(periods)
        assert gc.get_innermost().grating_period == max(periods) and gc.get_outermost().grating_period == min(periods)
        built[name] = gc
    ours, theirs = built["ours"], built["theirs"]
    for kw in (dict(angle_in_air=27.3 * degree), dict(grating_period=ours.grating_list[2].grating_period),
               dict(grating_period=1.005 * ours.grating_list[-1].grating_period),
               dict(lateral_period=900 * nm * math.tan(33 * degree))):
        a, b = ours.get_one(**kw), theirs.get_one(**kw)
        assert a.grating_period == b.grating_period and a.lateral_period == b.lateral_period
        assert np.array_equal(a.xyrra_list, b.xyrra_list)
    for mod in (grating, ref["grating"]):
        bad = [mod.Grating(lateral_period=700 * nm, cyl_height=550 * nm, grating_period=p * nm) for p in (1000, 1500)]
        with pytest.raises(AssertionError):          # lateral_period / tan(angle) not constant in a round collection
            mod.GratingCollection(target_wavelength=580 * nm, lateral_period=900 * nm, lens_type="round", grating_list=bad)
        with pytest.raises(AssertionError):
            mod.GratingCollection(target_wavelength=580 * nm, lateral_period=900 * nm, lens_type="square")


def test_hexgridset_surface(ref):
    """lens_center.py:25-57, :175-226: default grating list, pick_from_phase, the 'characterize() first' error."""
    ours = lens_center.HexGridSet(sep=320 * nm, cyl_height=550 * nm, n_glass=1.459, n_tio2=2.372, num_entries=7)
    theirs = ref["lens_center"].HexGridSet(sep=320 * nm, cyl_height=550 * nm, n_glass=1.459, n_tio2=2.372, num_entries=7)
    assert len(ours.grating_list) == len(theirs.grating_list) == 7
    for a, b in zip(ours.grating_list, theirs.grating_list):
        assert a.grating_period == b.grating_period and a.lateral_period == b.lateral_period
        assert np.array_equal(a.xyrra_list, b.xyrra_list)
    for obj in (ours, theirs):
        with pytest.raises(ValueError) as e:
            obj.build_interpolators()
        assert e.value.args == ('Need to run characterize() first',)
    amps = 0.8 * np.exp(2j * np.pi * np.arange(7) / 7 + 0.3j)
    ours.x_amp_list, theirs.x_amp_list = amps.copy(), amps.copy()
    for ph in np.linspace(-7, 7, 57):
        assert ours.pick_from_phase(ph) == theirs.pick_from_phase(ph)


def test_fft_bin_direction_cosines_are_the_reference_arithmetic():
    """nearfield_farfield.py:35-39, un-shifted: bit-identical (the evanescent NaN mask hinges on the last bit)."""
    for num, spacing, wl, n in ((128, 241.8e-9, 532e-9, 1.4607), (675, 263.6e-9, 580e-9, 1.459), (9, 200e-9, 450e-9, 1.466)):
        ux = np.arange(num) * (wl / n) / (spacing * num)
        ux[ux > ux.max() / 2] -= (wl / n) / spacing
        assert np.array_equal(farfield.fft_bin_direction_cosines(num, spacing, wl, n), ux)


def test_farfield_from_nearfield_rejects_what_the_reference_rejects(ref):
    """nearfield_farfield.py:22-30: AssertionError on shape mismatch, non-uniform axes, spacing >= wavelength/2,
    descending axes -- raised before any device work, so the same calls fail the same way without a GPU."""
    wl, n = 532e-9, 1.4607
    x = np.arange(16) * 200e-9
    y = np.arange(12) * 210e-9
    F = np.zeros((16, 12), dtype=complex)
    bad_calls = [
        (F, F, F, F[:, :11], x, y),                                   # shapes differ
        (F, F, F, F, x[:15], y),                                      # axis length != array shape
        (F, F, F, F, np.r_[x[:-1], x[-1] + 1e-12 * 3e4], y),          # not uniform (1e-9 relative, :29-30)
        (F, F, F, F, x * 1.4, y),                                     # 280 nm >= wavelength / 2
        (F, F, F, F, x[::-1].copy(), y),                              # descending
        (F, F, F, F, x, y[::-1].copy()),
    ]
    for args in bad_calls:
        with pytest.raises(AssertionError):
            farfield.farfield_from_nearfield(*args, wl, n)
        with pytest.raises(AssertionError):
            ref["nearfield_farfield"].farfield_from_nearfield(*args, wl, n)
        with pytest.raises(AssertionError):
            farfield.farfield_from_fields(*args, wl, n)


def test_build_nearfield_rejects_what_the_reference_rejects(ref):
    """nearfield.py:84-85 (source below the lens, source_pol in 'x','y','z'), :106-109 (uniform ascending axes finer than
    wavelength/2) and :224 (no z-polarised plane wave): AssertionError, raised before any device work."""
    from metalens_b200 import design
    x = np.linspace(-11.5 * um, 11.5 * um, 96)
    for g_mod, lc_mod, nf_mod, mk in ((grating, lens_center, nearfield, design.make_design),
                                      (ref["grating"], ref["lens_center"], ref["nearfield"],
                                       ref["design_collimator"].make_design)):
        spec = synth_lens.SMALL_LENS
        collections, hgs = synth_lens.make_library(g_mod, lc_mod, spec)
        periphery, center = mk(collections, spec["source_distance"], spec["radius"], hgs)[:2]
        args = (580 * nm, periphery, center, hgs)
        bad = [((0, 0, 10 * um, "x") + args, {}),                                   # source above the lens plane
               ((0, 0, -10 * um, "w") + args, {}),                                  # unknown polarisation
               ((0, 0, -10 * um, "x") + args, dict(x_pts=x * 3, y_pts=x)),          # 726 nm steps >= wavelength / 2
               ((0, 0, -10 * um, "x") + args, dict(x_pts=x, y_pts=x[::-1].copy())),  # descending
               ((0, 0, -10 * um, "x") + args, dict(x_pts=np.r_[x[:-1], x[-1] + 1e-12], y_pts=x)),   # not uniform
               ((0, 0, -float("inf"), "z") + args, dict(x_pts=x, y_pts=x))]         # z-polarised plane wave
        for a, kw in bad:
            with pytest.raises(AssertionError):
                nf_mod.build_nearfield(*a, **kw)
