"""GPU: lens layout on the device -- the hex-lattice centre laid out and binned by CUDA kernels equals the
host mirror of design_collimator.design_center row for row, and the assembly kernel gives the same fields from it."""
import os
import time

import numpy as np
import pytest

import synth_lens
from test_oracle_nearfield import periphery_from

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("radius,f", [(11.5e-6, 14.3e-6), (88.9e-6, 216.5e-6), (300e-6, 420e-6)])
def test_device_centre_equals_host_rows(radius, f):
    from metalens_b200 import design, grating, lens_center
    collections, hgs = synth_lens.make_library(grating, lens_center, synth_lens.SMALL_LENS)
    host = design.design_center(hgs, f, radius)
    dev = design.design_center_device(hgs, f, radius)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev = design.design_center_device(hgs, f, radius)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    got = dev.cpu().numpy()
    assert got.shape == host.shape
    assert np.array_equal(got[:, :2], host[:, :2])                       # same lattice points, same order, bit for bit
    agree = got[:, 2] == host[:, 2]
    # pick_from_phase is an argmax over 20 sines: a different last bit of sin/cos can only matter at an exact tie
    assert agree.mean() > 0.99999, agree.mean()
    print("device centre: %d cells in %.2f ms" % (got.shape[0], dt * 1e3))


def test_assembly_from_device_cells_equals_host_cells(golden_dir):
    """Same fields from the device-made, device-binned cells as from the host path (the bin order differs, the nearest
    cell does not), and the reference tie resolution works on top of it (148 exact ties of the odd on-axis grid)."""
    from metalens_b200 import grating, lens_center
    from metalens_b200.design import make_design
    from metalens_b200.nearfield import NearfieldPlan, build_nearfield
    g = np.load(os.path.join(golden_dir, "nearfield_mid_y_onaxis.npz"))
    spec = synth_lens.MID_LENS
    collections, hgs = synth_lens.make_library(grating, lens_center, spec)
    periph_h, center_h, _ = make_design(collections, spec["source_distance"], spec["radius"], hgs)
    periph_d, center_d, _ = make_design(collections, spec["source_distance"], spec["radius"], hgs, device="cuda")
    assert center_d.is_cuda and np.array_equal(center_d.cpu().numpy()[:, :2], center_h[:, :2])
    sx, sy, sz = g["source"]
    a = NearfieldPlan(580e-9, periph_h, center_h, hgs).run(sx, sy, sz, "y", g["x_pts"], g["y_pts"])[0].clone()
    b = NearfieldPlan(580e-9, periph_d, center_d, hgs).run(sx, sy, sz, "y", g["x_pts"], g["y_pts"])[0]
    same = (a.view(torch.float32) == b.view(torch.float32)).all(dim=-1)
    assert float(same.float().mean()) > 0.99999                           # up to cells whose argmax was an exact tie
    res = build_nearfield(sx, sy, sz, "y", 580e-9, periph_d, center_d, hgs, x_pts=g["x_pts"], y_pts=g["y_pts"])
    idx = g["index"]
    for k, key in enumerate(("Ex", "Ey", "Hx", "Hy")):
        scale = float(g["scale_E"] if k < 2 else g["scale_H"])
        assert np.abs(res[k].ravel()[idx] - g[key]).max() / scale < 1e-9, key
