"""Lens layout: host-side mirror of the layout half of the reference's
``design_collimator.py`` (SURVEY component 5, "next" row N2), vectorised with numpy.

``make_design`` returns the two objects ``build_nearfield`` consumes --
``lens_periphery_summary`` (dict of per-ring arrays, design_collimator.py:221-227) and
``lens_center_summary`` ((n,3) rows x, y, HexGridSet index; :124-137) -- in the same order and
with the same arithmetic as the reference, so a design made here equals one made there.
The CAD exporters (make_dxf / make_svg, :317-381) are out of scope.

Like the reference, the design wavelength, the hex pitch and the refractive index between
source and lens are module-level settings (design_collimator.py:34-54).
"""
import math

import numpy as np

from .units import nm, um

pi = math.pi

pitch = 320 * nm            # pillar centre-to-centre separation = lateral period of the hex cells
wavelength = 580 * nm       # design wavelength in vacuum
refractive_index = 1        # medium between the source and the lens


def target_phase(x, source_distance):
    """Phase the lens must impose at distance x from its centre (design_collimator.py:57-60)."""
    k = 2 * pi * refractive_index / wavelength
    return (-k * (np.sqrt(source_distance ** 2 + np.asarray(x, dtype=float) ** 2) - source_distance)) % (2 * pi)


def target_phase_zeros(radius, source_distance):
    """Radii where the target phase wraps, out to the first one >= radius (:62-70)."""
    k = 2 * pi * refractive_index / wavelength
    zeros = []
    order = 0
    while not zeros or zeros[-1] < radius:
        zeros.append((((2 * pi * order) / k + source_distance) ** 2 - source_distance ** 2) ** 0.5)
        order += 1
    return zeros


def hexagonal_grid(n, radius, fourfold_symmetry=True):
    """(x,y) of a hexagonal lattice with nearest-neighbour distance n inside x^2+y^2 < radius^2,
    in the reference's order: n2 (column) outer, n1 inner (design_collimator.py:74-118)."""
    if fourfold_symmetry:
        corners = [(0, 0), (radius, 0), (0, radius), (radius, radius)]
    else:
        corners = [(radius, radius), (radius, -radius), (-radius, radius), (-radius, -radius)]
    n1c = [y / n - x / (n * 3 ** 0.5) for x, y in corners]
    n2c = [2 * x / (n * 3 ** 0.5) for x, y in corners]
    n1 = np.arange(int(min(n1c)) - 2, int(max(n1c)) + 3)
    n2 = np.arange(int(min(n2c)) - 2, int(max(n2c)) + 3)
    x = (n * n2 * 3 ** 0.5 / 2)[:, None] + 0.0 * n1[None, :]
    y = n * (n1[None, :] + n2[:, None] / 2)
    keep = x ** 2 + y ** 2 < radius ** 2
    if fourfold_symmetry:
        keep &= (x >= 0) & (y >= 0)
    return np.stack((x[keep], y[keep]), axis=1)


def design_center(hgs, source_distance, radius):
    """Hex-lattice centre: for every cell the HexGridSet entry whose phase is closest to the target
    (+pi, design_collimator.py:130-135).  Returns rows [x, y, index] (:120-137)."""
    xy = hexagonal_grid(pitch, radius, fourfold_symmetry=False)
    if not hasattr(hgs, 'x_amp_list'):
        raise ValueError('Need to run characterize() first')                   # lens_center.py:178-179
    phase = target_phase((xy[:, 0] ** 2 + xy[:, 1] ** 2) ** 0.5, source_distance) + pi
    fom = (np.asarray(hgs.x_amp_list)[None, :] * np.exp(-1j * phase)[:, None]).imag    # pick_from_phase
    return np.column_stack((xy, np.argmax(fom, axis=1).astype(float)))


def _hex_ranges(n, radius):
    """Candidate index ranges of hexagonal_grid(fourfold_symmetry=False) (design_collimator.py:85-103)."""
    corners = [(radius, radius), (radius, -radius), (-radius, radius), (-radius, -radius)]
    n1c = [y / n - x / (n * 3 ** 0.5) for x, y in corners]
    n2c = [2 * x / (n * 3 ** 0.5) for x, y in corners]
    return (int(min(n1c)) - 2, int(max(n1c)) + 2), (int(min(n2c)) - 2, int(max(n2c)) + 2)


def design_center_device(hgs, source_distance, radius, device=None):
    """design_center() on the GPU: the same rows [x, y, index] in the same order, as a CUDA float64 (n, 3)
    tensor that NearfieldPlan consumes directly -- the ~10^6-cell lattice never exists on the host.  Three launches
    (count per lattice column, scan, fill) through the C-ABI (csrc/design.cu)."""
    import ctypes as C
    import torch
    from . import _lib
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.MetalensB200Error("metalens_b200 needs a CUDA device (no CPU fallback)")
    if not hasattr(hgs, 'x_amp_list'):
        raise ValueError('Need to run characterize() first')                   # lens_center.py:178-179
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    (n1_lo, n1_hi), (n2_lo, n2_hi) = _hex_ranges(pitch, radius)
    cols = n2_hi - n2_lo + 1
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    counts = torch.empty(cols, dtype=torch.int32, device=dev)
    _lib.check(lib.mlb_hex_count(pitch, radius, n1_lo, n1_hi, n2_lo, n2_hi, counts.data_ptr(), stream), "mlb_hex_count")
    ends = torch.cumsum(counts, dim=0, dtype=torch.int64)
    offsets = ends - counts
    n = int(ends[-1].item())
    cells = torch.empty((n, 3), dtype=torch.float64, device=dev)
    if n:
        amp = torch.from_numpy(np.ascontiguousarray(np.asarray(hgs.x_amp_list, dtype=np.complex128)).view(np.float64)).to(dev)
        _lib.check(lib.mlb_hex_fill(pitch, radius, n1_lo, n1_hi, n2_lo, n2_hi, offsets.data_ptr(), wavelength,
                                    float(refractive_index), source_distance, amp.data_ptr(), amp.numel() // 2,
                                    cells.data_ptr(), stream), "mlb_hex_fill")
    return cells


def design_periphery(collections, source_distance, radius):
    """Ring layout of the grating periphery (design_collimator.py:148-228): one ring per 2-pi zone of
    the target phase beyond the switch radius, each assigned to the collection whose angle band
    contains it.  ``collections`` = [[(phi_start, phi_end), GratingCollection], ...]."""
    assert len(collections) > 0
    for a, b in zip(collections[:-1], collections[1:]):
        assert a[0][1] == b[0][0]
    assert all(band[0] < band[1] for band, _ in collections)
    switch_r = source_distance * math.tan(collections[0][0][0])
    zeros = [x for x in target_phase_zeros(radius + 2 * um, source_distance) if x > switch_r]
    if len(zeros) <= 1:
        raise ValueError('Periphery is too small for even one ring')
    rows = []          # (r_center, period, collection index, copies around the circle)
    ci, zi = 0, 0
    while True:
        r_in, r_out = zeros[zi], zeros[zi + 1]
        r_mid = (r_out + r_in) / 2
        if collections[ci][0][1] < math.atan(r_mid / source_distance):
            ci += 1
            if ci >= len(collections):
                raise ValueError('radius is too big for provided collections')
            continue
        gc = collections[ci][1]
        # 2 pi f / (lateral_period / tan(angle)) == 2 pi r / lateral_period  (:176-186)
        rows.append((r_mid, r_out - r_in, ci, int(round(2 * pi * source_distance / gc.lateral_period))))
        if r_out > radius:
            break
        zi += 1
    r_center = np.array([r[0] for r in rows])
    period = np.array([r[1] for r in rows])
    return {'gratingcollection_list': [c[1] for c in collections],
            'r_center_list': r_center,
            'r_min_list': r_center - 0.5 * period,
            'r_max_list': r_center + 0.5 * period,
            'grating_period_list': period,
            'gratingcollection_index_here_list': np.array([r[2] for r in rows]),
            'num_around_circle_list': np.array([r[3] for r in rows])}


def make_design(collections, source_distance, radius, hgs, make_xyrra_list=False, device=None):
    """Full round-lens design (design_collimator.py:273-313): returns
    (lens_periphery_summary, lens_center_summary, r_for_switch).  device='cuda' (or a torch device) lays the hex centre
    out on the GPU: lens_center_summary is then a CUDA (n, 3) tensor with the same rows, which NearfieldPlan /
    build_nearfield take as is (the ring arrays are O(rings) and stay on the host)."""
    if make_xyrra_list:
        raise NotImplementedError("pillar lists for the CAD exporters are outside this engine")
    if len(collections) > 0:
        for _, gc in collections:
            assert gc.lens_type == 'round'
            for g in gc.grating_list:
                assert g.n_tio2 == hgs.n_tio2 and g.n_glass == hgs.n_glass and g.cyl_height == hgs.cyl_height
        periphery = design_periphery(collections, source_distance, radius)
        r_for_switch = periphery['r_min_list'][0]
        assert r_for_switch < radius
    else:
        periphery, r_for_switch = None, radius
    if device is not None:
        center = design_center_device(hgs, source_distance, r_for_switch - 300 * nm,
                                      None if device in ("cuda", True) else device)
    else:
        center = design_center(hgs, source_distance, r_for_switch - 300 * nm)
    return periphery, center, r_for_switch
