"""GPU: the drop-in exactly as INTEGRATION.md section 1 prescribes it, against the LIVE unmodified reference.

baseline/_ref/ holds the reference's own modules (installed by __graft_entry__.build() where /root/reference exists;
git-ignored, shipped to the GPU box with the snapshot).  The test builds the synthetic lens with the REFERENCE's classes
and make_design, runs the reference's build_nearfield / farfield_from_nearfield, then patches this package's functions into
the reference modules (the three assignments of INTEGRATION.md) and runs the same calls on the same reference objects.
Skipped where baseline/_ref is not installed; the committed fixtures (tests/golden) pin the same thing without it.
"""
import contextlib
import io

import numpy as np
import pytest

import synth_lens

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from baseline import install_ref
    mods = install_ref.load()
    if mods is None:
        pytest.skip("baseline/_ref not installed")
    return mods


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("source,pol", [((0.0, 0.0), "x"), ((0.6e-6, -0.3e-6), "z")])
def test_patched_reference_modules_give_the_reference_results(ref, source, pol):
    import metalens_b200.farfield as ff
    import metalens_b200.nearfield as nf
    G, LC, DC, NF, NFF = (ref[k] for k in ("grating", "lens_center", "design_collimator", "nearfield", "nearfield_farfield"))
    spec = synth_lens.SMALL_LENS
    collections, hgs = synth_lens.make_library(G, LC, spec)                    # the reference's own classes and tables
    periph, center, _ = _quiet(DC.make_design, collections, spec["source_distance"], spec["radius"], hgs)
    args = dict(source_x=source[0], source_y=source[1], source_z=-spec["source_distance"], source_pol=pol,
                wavelength=580e-9, lens_periphery_summary=periph, lens_center_summary=center, hexgridset=hgs)
    want = _quiet(NF.build_nearfield, **args)                                  # unmodified reference, CPU
    fft = [np.fft.fft2(np.fft.fftshift(a)) for a in want[:4]]                  # nearfield_farfield.py:18-20
    want_ff = _quiet(NFF.farfield_from_nearfield, *fft, want[4], want[5], 580e-9, want[7])
    saved = (NF.build_nearfield, NF.build_nearfield_big, NFF.farfield_from_nearfield)
    try:
        NF.build_nearfield = nf.build_nearfield
        NF.build_nearfield_big = nf.build_nearfield_big
        NFF.farfield_from_nearfield = ff.farfield_from_nearfield
        got = NF.build_nearfield(**args)
        got_big = NF.build_nearfield_big(x_pts=want[4], y_pts=want[5], **args)
        got_ff = NFF.farfield_from_nearfield(*fft, want[4], want[5], 580e-9, want[7])
    finally:
        NF.build_nearfield, NF.build_nearfield_big, NFF.farfield_from_nearfield = saved
    escale = max(np.abs(want[0]).max(), np.abs(want[1]).max())
    hscale = max(np.abs(want[2]).max(), np.abs(want[3]).max())
    for k in range(4):
        scale = escale if k < 2 else hscale
        assert np.abs(got[k] - want[k]).max() / scale < 1e-9
        assert np.abs(got_big[k] - want[k]).max() / scale < 1e-9
    np.testing.assert_array_equal(got[4], want[4])
    np.testing.assert_array_equal(got[5], want[5])
    assert abs(got[6] - want[6]) <= 1e-11 * abs(want[6]) and got[7] == want[7]
    # far field: same tuple, float64 P at 1e-11 of the peak, identical NaN mask and axes
    P, total, ux, uy, dux, duy = got_ff
    assert np.array_equal(np.isnan(P), np.isnan(want_ff[0]))
    fin = np.isfinite(want_ff[0])
    assert np.abs(P - want_ff[0])[fin].max() / want_ff[0][fin].max() < 1e-11
    assert abs(total - want_ff[1]) <= 1e-11 * abs(want_ff[1])
    np.testing.assert_array_equal(ux, want_ff[2])
    np.testing.assert_array_equal(uy, want_ff[3])
    assert dux == want_ff[4] and duy == want_ff[5]
