"""Amplitude tables: the input contract of the aperture-field assembly.

The reference wraps every ``(wavelength_nm, (ox,oy), pol, amp)`` table in a scipy
``RegularGridInterpolator`` (grating.py:1227-1229, lens_center.py:222-223) and
``build_nearfield`` calls those objects on ``(n,3)`` point arrays (nearfield.py:310-311).
Here the same tables are kept as plain arrays (:class:`AmplitudeTable`, same ``.grid`` /
``.values`` attributes as the scipy object, so either kind can be handed to
``build_nearfield``), and flattened into one device-resident pack per collection
(:class:`TablePack`) that the CUDA kernels gather from.  Calling an
:class:`AmplitudeTable` evaluates it on the GPU (``mlb_table_eval``).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class AmplitudeTable:
    """One complex amplitude on a rectilinear (ux, uy, third-axis) grid.

    ``table(points)`` with points of shape (n,3) returns the trilinear interpolant,
    like the reference's RegularGridInterpolator objects with their default
    ``method='linear', bounds_error=True``: out-of-range points raise ValueError.
    """

    def __init__(self, grid, values):
        self.grid = tuple(np.ascontiguousarray(g, dtype=np.float64) for g in grid)
        self.values = np.ascontiguousarray(values, dtype=np.complex128)
        assert self.values.shape == tuple(g.size for g in self.grid)
        for g in self.grid:
            assert g.size >= 2 and np.all(np.diff(g) > 0), "grid axes must be strictly ascending"
        self._dev = None

    def _device_arrays(self):
        if self._dev is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            axes = torch.from_numpy(np.concatenate(self.grid)).to(dev)
            vals = torch.from_numpy(self.values.reshape(-1)).to(dev)
            self._dev = (axes, vals)
        return self._dev

    def __call__(self, xi):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MetalensB200Error("metalens_b200 needs a CUDA device (no CPU fallback)")
        pts = np.ascontiguousarray(xi, dtype=np.float64)
        lead = pts.shape[:-1]
        pts = pts.reshape(-1, 3)
        for d in range(3):
            lo, hi = self.grid[d][0], self.grid[d][-1]
            if pts.size and (pts[:, d].min() < lo or pts[:, d].max() > hi):
                raise ValueError("One of the requested xi is out of bounds in dimension %d" % d)
        axes, vals = self._device_arrays()
        dpts = torch.from_numpy(pts).to(axes.device)
        out = torch.empty(pts.shape[0], dtype=torch.complex128, device=axes.device)
        n = [g.size for g in self.grid]
        rc = lib.mlb_table_eval(axes.data_ptr(), n[0], n[1], n[2], vals.data_ptr(), dpts.data_ptr(),
                                pts.shape[0], out.data_ptr(), _stream_ptr())
        _lib.check(rc, "mlb_table_eval")
        return out.cpu().numpy().reshape(lead)


def orders_of(grating_list):
    """The (ox,oy) set of nearfield.py:264 / :390, built with the same expression so that
    its iteration order matches the reference's within one process."""
    return {(e['ox'], e['oy']) for g in grating_list for e in g.data}


class TablePack:
    """Flattened tables of ONE GratingCollection / HexGridSet for the assembly kernel.

    Layout (all float64 / complex128, contiguous):
      axes   : ux nodes | uy nodes | third-axis nodes
      values : [order][iu][iv][ig][slot]  with slot = 2*pol + amp,
               pol: 0='x', 1='y'; amp: 0='ampfy', 1='ampfx'
               (the 4 complex values one interpolation corner needs are 64 contiguous bytes)
      orders : int32 [n_orders][2] in the reference's iteration order
    """
    SLOTS = (('x', 'ampfy'), ('x', 'ampfx'), ('y', 'ampfy'), ('y', 'ampfx'))

    def __init__(self, owner, wavelength_in_nm):
        self.orders = list(orders_of(owner.grating_list))
        grid = None
        blocks = []
        for (ox, oy) in self.orders:
            slot_vals = []
            for pol, amp in self.SLOTS:
                f = owner.interpolators[(wavelength_in_nm, (ox, oy), pol, amp)]
                g = tuple(np.asarray(a, dtype=np.float64) for a in f.grid)
                if grid is None:
                    grid = g
                else:
                    assert all(np.array_equal(a, b) for a, b in zip(grid, g)), "tables of one collection share axes"
                slot_vals.append(np.asarray(f.values, dtype=np.complex128))
            blocks.append(np.stack(slot_vals, axis=0))             # [slot][iu][iv][ig]: the arrays as the reference holds them
        self.grid = grid
        self.n = tuple(a.size for a in grid)
        # raw: what is uploaded; the interleaved layout [order][iu][iv][ig][slot] the kernel gathers from is made on the
        # device by mlb_table_pack_build (NearfieldPlan).  `values` is the same interleaving on the host, for tests / the oracle.
        self.raw = np.ascontiguousarray(np.stack(blocks, axis=0)) if blocks else np.zeros((0, 4) + tuple(a.size for a in grid), complex)
        self.values = np.ascontiguousarray(np.moveaxis(self.raw, 1, -1))   # [order][iu][iv][ig][slot]
        self.bounds = tuple(float(b) for b in owner.interpolator_bounds)
        self.axes = np.concatenate(grid)
        # uniformly spaced ux / uy axes (what characterize() produces): the kernel may locate cells arithmetically
        def uniform(a):
            d = np.diff(a)
            return bool(a.size >= 2 and np.all(np.abs(d - d[0]) <= 1e-12 * abs(d[0])))
        self.uniform01 = uniform(grid[0]) and uniform(grid[1])
        self.u_first = (float(grid[0][0]), float(grid[1][0]))
        self.u_inv_step = ((grid[0].size - 1) / float(grid[0][-1] - grid[0][0]),
                           (grid[1].size - 1) / float(grid[1][-1] - grid[1][0]))
        self.order_array = np.asarray(self.orders, dtype=np.int32).reshape(-1, 2)
        # dense (ox,oy) -> order index map so the kernel can visit just the orders that may propagate
        self.order_radius = int(np.abs(self.order_array).max()) if len(self.orders) else 0
        w = 2 * self.order_radius + 1
        self.order_map = np.full(w * w, -1, dtype=np.int32)
        for k, (ox, oy) in enumerate(self.orders):
            self.order_map[(ox + self.order_radius) * w + (oy + self.order_radius)] = k


# ---------------------------------------------------------------------------------------------
# Packed on-disk library format.  The reference persists a characterised collection as a
# multi-megabyte Python repr() and rebuilds the interpolators from the row dicts on every start
# (README.md:29-34, grating.py:1186-1232).  Here the dense tables themselves are stored (one
# compressed .npz, versioned) and come back as ready interpolators -- no row loop at load time.
LIBRARY_FORMAT_VERSION = 1


def save_library(path, owner):
    """Write a GratingCollection or HexGridSet with built interpolators to `path` (.npz)."""
    is_hgs = hasattr(owner, "sep")
    keys = sorted(owner.interpolators, key=repr)
    first = owner.interpolators[keys[0]]
    gl = owner.grating_list
    arrays = dict(
        format_version=np.int64(LIBRARY_FORMAT_VERSION),
        kind=np.str_("hexgridset" if is_hgs else "gratingcollection"),
        axis0=np.asarray(first.grid[0], float), axis1=np.asarray(first.grid[1], float),
        axis2=np.asarray(first.grid[2], float),
        key_wavelength=np.array([k[0] for k in keys], np.int64),
        key_order=np.array([k[1] for k in keys], np.int64),
        key_pol=np.array([k[2] for k in keys]), key_amp=np.array([k[3] for k in keys]),
        values=np.stack([np.asarray(owner.interpolators[k].values, np.complex128) for k in keys]),
        bounds=np.asarray(owner.interpolator_bounds, float),
        orders=np.asarray(list(orders_of(gl)), np.int64).reshape(-1, 2),
        g_grating_period=np.array([g.grating_period for g in gl], float),
        g_lateral_period=np.array([g.lateral_period for g in gl], float),
        g_cyl_height=np.array([g.cyl_height for g in gl], float),
        g_n_glass=np.array([g.n_glass for g in gl], float), g_n_tio2=np.array([g.n_tio2 for g in gl], float))
    if is_hgs:
        arrays.update(sep=np.float64(owner.sep), cyl_height=np.float64(owner.cyl_height),
                      n_glass=np.float64(owner.n_glass), n_tio2=np.float64(owner.n_tio2),
                      x_amp_list=np.asarray(owner.x_amp_list, np.complex128))
    else:
        arrays.update(target_wavelength=np.float64(owner.target_wavelength),
                      lateral_period=np.float64(owner.lateral_period), lens_type=np.str_(owner.lens_type))
    np.savez_compressed(path, **arrays)


def load_library(path):
    """Inverse of save_library: returns a GratingCollection / HexGridSet whose `.interpolators` and
    `.interpolator_bounds` are ready.  `Grating.data` holds only the order list build_nearfield needs
    (one stub row per order); the characterisation rows themselves are not stored."""
    from . import grating as G
    from . import lens_center as LC
    z = np.load(path, allow_pickle=False)
    if int(z["format_version"]) != LIBRARY_FORMAT_VERSION:
        raise ValueError("unsupported library format version %d" % int(z["format_version"]))
    gratings = []
    for gp, lp, ch, ng, nt in zip(z["g_grating_period"], z["g_lateral_period"], z["g_cyl_height"], z["g_n_glass"],
                                  z["g_n_tio2"]):
        def num(v):
            return int(v) if float(v) == int(v) else float(v)
        g = G.Grating(lateral_period=float(lp), cyl_height=float(ch), grating_period=float(gp), n_glass=num(ng),
                      n_tio2=num(nt))
        g.data = []
        gratings.append(g)
    gratings[0].data = [{"ox": int(ox), "oy": int(oy)} for ox, oy in z["orders"]]
    if str(z["kind"]) == "hexgridset":
        owner = LC.HexGridSet(sep=float(z["sep"]), cyl_height=float(z["cyl_height"]), n_glass=num(z["n_glass"]),
                              n_tio2=num(z["n_tio2"]), grating_list=gratings, x_amp_list=z["x_amp_list"])
    else:
        owner = G.GratingCollection.__new__(G.GratingCollection)
        owner.target_wavelength = float(z["target_wavelength"])
        owner.lateral_period = float(z["lateral_period"])
        owner.target_kvac = 2 * np.pi / owner.target_wavelength
        owner.lens_type = str(z["lens_type"])
        owner.grating_list = gratings
    grid = (z["axis0"], z["axis1"], z["axis2"])
    owner.interpolators = {}
    for i in range(z["values"].shape[0]):
        key = (int(z["key_wavelength"][i]), (int(z["key_order"][i][0]), int(z["key_order"][i][1])),
               str(z["key_pol"][i]), str(z["key_amp"][i]))
        owner.interpolators[key] = AmplitudeTable(grid, z["values"][i])
    b = z["bounds"]
    owner.interpolator_bounds = tuple(float(v) for v in b)
    return owner
