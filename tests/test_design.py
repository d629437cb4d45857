"""CPU: metalens_b200.design.make_design reproduces the reference's make_design outputs stored in
the near-field fixtures (ring arrays and centre cells, same order)."""
import os

import numpy as np
import pytest

import synth_lens
from metalens_b200 import design, grating, lens_center


@pytest.mark.parametrize("name,spec", [("small_x_onaxis", synth_lens.SMALL_LENS), ("plane_x", synth_lens.PLANE_LENS)])
def test_make_design_matches_reference(name, spec, golden_dir):
    g = np.load(os.path.join(golden_dir, "nearfield_%s.npz" % name))
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph, center, r_switch = design.make_design(collections, spec["source_distance"], spec["radius"], hgs)
    for k in ("r_center_list", "r_min_list", "r_max_list", "grating_period_list"):
        np.testing.assert_allclose(periph[k], g["periph_" + k], rtol=1e-15, atol=0)
    for k in ("gratingcollection_index_here_list", "num_around_circle_list"):
        np.testing.assert_array_equal(periph[k], g["periph_" + k])
    assert r_switch == pytest.approx(float(g["r_switch"]), rel=1e-15)
    assert center.shape == g["center"].shape
    np.testing.assert_allclose(center[:, :2], g["center"][:, :2], rtol=1e-14, atol=1e-22)
    np.testing.assert_array_equal(center[:, 2], g["center"][:, 2])


def test_design_errors():
    collections, hgs = synth_lens.make_library(grating, lens_center, synth_lens.SMALL_LENS)
    with pytest.raises(ValueError):
        design.make_design(collections, 14.3e-6, 40e-6, hgs)        # radius beyond the last band
    with pytest.raises(ValueError):
        design.design_periphery(collections, 14.3e-6, 1e-6)         # no room for a ring
    bare = lens_center.HexGridSet(sep=320e-9, cyl_height=550e-9, grating_list=hgs.grating_list)
    with pytest.raises(ValueError):
        design.design_center(bare, 14.3e-6, 3e-6)
