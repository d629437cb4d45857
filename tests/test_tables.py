"""CPU: the table builders produce the reference's tables bit for bit, and TablePack -- the flattening the
assembly kernel reads -- accepts the reference's own scipy RegularGridInterpolator objects."""
import os

import numpy as np
import pytest

import synth_lens
from metalens_b200 import grating, lens_center
from metalens_b200.tables import TablePack


def _owners():
    collections, hgs = synth_lens.make_library(grating, lens_center, synth_lens.SMALL_LENS)
    return [("gc0", collections[0][1]), ("gc1", collections[1][1]), ("hgs", hgs)]


def test_tables_equal_reference_fixture(golden_dir):
    """build_interpolators (grating.py:1186-1232, lens_center.py:188-226): same keys, same grids, same complex values,
    same bounds as the UNMODIFIED reference (tests/golden/make_tables_golden.py) -- bitwise."""
    g = np.load(os.path.join(golden_dir, "tables_small_lens.npz"))
    for name, owner in _owners():
        keys = sorted(owner.interpolators, key=repr)
        assert [repr(k) for k in keys] == list(g[name + "_keys"])
        vals = np.stack([owner.interpolators[k].values for k in keys])
        assert vals.dtype == np.complex128 and np.array_equal(vals, g[name + "_values"])
        for a in range(3):
            assert np.array_equal(np.asarray(owner.interpolators[keys[0]].grid[a], float), g["%s_grid%d" % (name, a)])
        assert np.array_equal(np.asarray(owner.interpolator_bounds, float), g[name + "_bounds"])


def test_table_pack_accepts_reference_objects():
    """The unmodified reference's classes (installed copy, baseline/_ref) build the same library; TablePack flattens
    their scipy interpolators into exactly the pack it makes from ours."""
    from baseline import install_ref
    ref = install_ref.load()
    if ref is None:
        pytest.skip("baseline/_ref not installed (run __graft_entry__.build() where /root/reference exists)")
    r_coll, r_hgs = synth_lens.make_library(ref["grating"], ref["lens_center"], synth_lens.SMALL_LENS)
    theirs = [r_coll[0][1], r_coll[1][1], r_hgs]
    for (name, ours), other in zip(_owners(), theirs):
        a, b = TablePack(ours, 580), TablePack(other, 580)
        assert sorted(a.orders) == sorted(b.orders)
        perm = [b.orders.index(o) for o in a.orders]
        assert np.array_equal(a.values, b.values[perm]) and np.array_equal(a.axes, b.axes)
        assert a.bounds == b.bounds and a.n == b.n and a.uniform01 == b.uniform01
