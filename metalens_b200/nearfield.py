"""Aperture-field assembly on the B200 (hot path B of SURVEY.md section 8).

Python host side of the reference's ``nearfield.py``: :func:`build_nearfield` and
:func:`build_nearfield_big` keep the reference signatures, return tuple, assertions and
``ValueError`` messages (nearfield.py:66-516); :func:`good_fft_number` is nearfield.py:30-36.
The per-sample work (ring lookup, incident dipole field, grating-frame rotation,
diffraction-order loop with trilinear table gathers, nearest hex cell, propagation phases)
is ONE fused CUDA kernel, ``mlb_nearfield_assemble`` (csrc/nearfield.cu).  The host only
packs the design (ring arrays, cell bins) and the amplitude tables into device arrays, once
per lens (:class:`NearfieldPlan`, cached).  No CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from . import grating as _grating
from .tables import TablePack
from .units import nm, c0, Z0

inf = float('inf')
MAX_PACKS = 12
STATS_PER_ORDER = 8


class _PackC(C.Structure):
    _fields_ = [("axes", C.c_void_p), ("values", C.c_void_p), ("values_f32", C.c_void_p), ("orders", C.c_void_p),
                ("order_map", C.c_void_p),
                ("n_ux", C.c_int), ("n_uy", C.c_int), ("n_g", C.c_int), ("n_orders", C.c_int),
                ("bounds", C.c_double * 6), ("stats_slot", C.c_int), ("order_radius", C.c_int),
                ("uniform01", C.c_int), ("_pad", C.c_int),
                ("u0_first", C.c_double), ("u0_inv_step", C.c_double), ("u1_first", C.c_double), ("u1_inv_step", C.c_double)]


class _LensC(C.Structure):
    _fields_ = [("x_pts", C.c_void_p), ("y_pts", C.c_void_p), ("nx", C.c_int), ("ny", C.c_int),
                ("ring_boundary", C.c_void_p), ("r_center", C.c_void_p), ("grating_period", C.c_void_p),
                ("num_around", C.c_void_p), ("gc_index", C.c_void_p), ("n_rings", C.c_int), ("n_packs", C.c_int),
                ("packs", _PackC * MAX_PACKS),
                ("cell_x", C.c_void_p), ("cell_y", C.c_void_p), ("cell_which", C.c_void_p), ("cell_orig", C.c_void_p),
                ("bin_start", C.c_void_p), ("n_cells", C.c_int), ("nbx", C.c_int), ("nby", C.c_int), ("_pad", C.c_int),
                ("bin_x0", C.c_double), ("bin_y0", C.c_double), ("bin_size", C.c_double),
                ("hex", _PackC), ("hex_x_period", C.c_double), ("hex_y_period", C.c_double),
                ("source_x", C.c_double), ("source_y", C.c_double), ("source_z", C.c_double),
                ("plane_wave", C.c_int), ("source_pol", C.c_int),
                ("wavelength", C.c_double), ("n_glass", C.c_double), ("dipole_moment", C.c_double),
                ("c0", C.c_double), ("Z0", C.c_double),
                ("ring_aux", C.c_void_p), ("ring_aux_f32", C.c_void_p), ("ring_lut", C.c_void_p),
                ("n_lut", C.c_int), ("_pad2", C.c_int), ("lut_r_max", C.c_double),
                ("ring_tables", C.c_void_p), ("ring_table_stride", C.c_longlong)]


def good_fft_number(goal):
    """Smallest number >= goal whose only prime factors are 2, 3, 5 (nearfield.py:30-36)."""
    assert goal < 1e5
    best = None
    a = 1
    while a < 2e5:
        b = a
        while b < 2e5:
            c = b
            while c < 2e5:
                if c >= goal and (best is None or c < best):
                    best = c
                c *= 5
            b *= 3
        a *= 2
    return best


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _enc_inf(sign):
    """int64 image of +/-inf under the kernel's order-preserving double encoding."""
    b = np.array([sign * np.inf]).view(np.int64)[0]
    return int(b) if b >= 0 else int(b ^ 0x7fffffffffffffff)


def _dec(v):
    v = np.asarray(v, dtype=np.int64)
    raw = np.where(v >= 0, v, v ^ np.int64(0x7fffffffffffffff))
    return raw.view(np.float64)


class NearfieldPlan:
    """Device-resident packing of one lens design + amplitude tables at one wavelength.

    Parameters are the reference's own objects: ``lens_periphery_summary`` (dict,
    design_collimator.py:221-227), ``lens_center_summary`` ((n,3) rows x, y, index;
    design_collimator.py:124-137) and a ``HexGridSet`` with interpolators built.
    The collections may be this package's classes or the reference's (anything with
    ``.grating_list[*].data/.n_glass``, ``.interpolators[key].grid/.values``,
    ``.interpolator_bounds``).
    """

    def __init__(self, wavelength, lens_periphery_summary, lens_center_summary, hexgridset, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MetalensB200Error("metalens_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        dev = self.device
        self.wavelength = float(wavelength)
        self.wavelength_in_nm = int(round(wavelength / nm))                       # nearfield.py:86
        P = lens_periphery_summary
        self.gcs = P['gratingcollection_list']                                   # :93
        r_min = np.asarray(P['r_min_list'], dtype=np.float64)                    # :87
        r_max = np.asarray(P['r_max_list'], dtype=np.float64)                    # :88
        self.lens_max_r = float(r_max[-1])                                       # :94
        assert len(self.gcs) <= MAX_PACKS, "too many GratingCollections for one kernel launch"
        n_glass = self.gcs[0].grating_list[0].n_glass                            # :111
        if n_glass == 0:
            n_glass = _grating.n_glass(self.wavelength_in_nm)                    # :113 (ValueError on unknown)
        self.n_glass = n_glass

        def up(a, dtype):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(dev)

        self.n_rings = r_min.size
        self._keep = dict(
            ring_boundary=up(np.hstack((r_min, self.lens_max_r)), np.float64),   # :125
            r_center=up(P['r_center_list'], np.float64),
            grating_period=up(P['grating_period_list'], np.float64),
            num_around=up(P['num_around_circle_list'], np.float64),
            gc_index=up(P['gratingcollection_index_here_list'], np.int32))

        # amplitude tables -> device packs (one per collection + the HexGridSet)
        self.packs = [TablePack(gc, self.wavelength_in_nm) for gc in self.gcs]
        self.hexgridset = hexgridset
        self.hex_pack = TablePack(hexgridset, self.wavelength_in_nm)
        self.hex_x_period = float(hexgridset.grating_list[0].grating_period)     # :391
        self.hex_y_period = float(hexgridset.grating_list[0].lateral_period)     # :392
        slot = 0
        self._pack_dev = []
        for p in self.packs + [self.hex_pack]:
            p.stats_slot = slot
            slot += len(p.orders)
            # upload the interpolators' arrays as they are; the gather layout (and its complex64 copy) is made on the device
            raw = up(p.raw.reshape(-1).view(np.float64), np.float64)
            vals = torch.empty(raw.numel(), dtype=torch.float64, device=dev)
            vals32 = torch.empty(raw.numel(), dtype=torch.float32, device=dev)
            _lib.check(self.lib.mlb_table_pack_build(raw.data_ptr(), len(p.orders), p.n[0], p.n[1], p.n[2], vals.data_ptr(),
                                               vals32.data_ptr(), _stream_ptr()), "mlb_table_pack_build")
            self._pack_dev.append((up(p.axes, np.float64), vals, vals32,
                                   up(p.order_array, np.int32), up(p.order_map, np.int32)))
        self.n_stats = slot

        # centre cells -> uniform bin grid (replaces the reference's cKDTree, nearfield.py:363).  About one cell per bin:
        # with the exact reach test of the kernel a sample visits its own bin and the ring of eight around it (~9 cells)
        self._cells_dev = None
        if torch.is_tensor(lens_center_summary) and lens_center_summary.is_cuda:
            self._bin_on_device(lens_center_summary.to(dev).reshape(-1, 3).contiguous(), up)
        else:
            cells = np.asarray(lens_center_summary.cpu() if torch.is_tensor(lens_center_summary) else lens_center_summary,
                               dtype=np.float64).reshape(-1, 3)
            self.n_cells = cells.shape[0]
            if self.n_cells:
                cx, cy = cells[:, 0], cells[:, 1]
                self._bin_geometry(cx.min(), cx.max(), cy.min(), cy.max())
                b = self.bin_size
                bx = np.minimum(np.floor((cx - self.bin_x0) / b).astype(np.int64), self.nbx - 1)
                by = np.minimum(np.floor((cy - self.bin_y0) / b).astype(np.int64), self.nby - 1)
                key = by * self.nbx + bx
                order = np.argsort(key, kind="stable")
                self._cells_host = cells[:, 0:2].copy()              # for the reference's tie choice (_resolve_ties)
                self._sorted_pos = np.empty(self.n_cells, dtype=np.int64)
                self._sorted_pos[order] = np.arange(self.n_cells)
                counts = np.bincount(key, minlength=self.nbx * self.nby)
                self._keep.update(
                    cell_x=up(cx[order], np.float64), cell_y=up(cy[order], np.float64),
                    cell_which=up(cells[order, 2].astype(np.int64), np.int32),        # .astype(int), :367
                    cell_orig=up(order, np.int32),
                    bin_start=up(np.concatenate(([0], np.cumsum(counts))), np.int32))
            else:
                self.bin_x0 = self.bin_y0 = 0.0
                self.bin_size = 1.0
                self.nbx = self.nby = 1

        # per-lens derived data, filled on the device by mlb_nearfield_prepare (ring records, ring bin table)
        self.n_lut = max(16, 4 * self.n_rings)
        self._keep.update(
            ring_aux=torch.zeros(self.n_rings * 8, dtype=torch.float64, device=dev),
            ring_aux_f32=torch.zeros(self.n_rings * 8, dtype=torch.float32, device=dev),
            ring_lut=torch.zeros(self.n_lut + 1, dtype=torch.int32, device=dev))
        self._keep['ring_tables'] = None
        self.ring_table_stride = 0
        L = self._desc(self._keep['ring_boundary'], self._keep['ring_boundary'], 0.0, 0.0, -1.0, 'x', 1.0)
        self.ring_table_stride = int(self.lib.mlb_nearfield_ring_table_floats(C.byref(L)))
        self._keep['ring_tables'] = torch.zeros(self.n_rings * self.ring_table_stride, dtype=torch.float32, device=dev)
        L = self._desc(self._keep['ring_boundary'], self._keep['ring_boundary'], 0.0, 0.0, -1.0, 'x', 1.0)
        _lib.check(self.lib.mlb_nearfield_prepare(C.byref(L), _stream_ptr()), "mlb_nearfield_prepare")

    # ------------------------------------------------------------------
    def _bin_geometry(self, x_min, x_max, y_min, y_max):
        span_x, span_y = float(x_max - x_min), float(y_max - y_min)
        area = max(span_x * span_y, 1e-300)
        b = math.sqrt(area / self.n_cells) if area > 1e-300 else 1.0
        b = max(b, 1e-12 * max(span_x, span_y, 1e-30))
        self.bin_x0, self.bin_y0, self.bin_size = float(x_min), float(y_min), float(b)
        self.nbx = int(math.floor(span_x / b)) + 1
        self.nby = int(math.floor(span_y / b)) + 1

    def _bin_on_device(self, cells, up):
        """Cells that already live on the GPU (design.design_center_device): bin them there (mlb_cells_bin: bin
        populations, scan, scatter) -- no host sort, no upload."""
        dev = self.device
        self.n_cells = int(cells.shape[0])
        self._cells_dev = cells
        if not self.n_cells:
            self.bin_x0 = self.bin_y0 = 0.0
            self.bin_size = 1.0
            self.nbx = self.nby = 1
            return
        lo = cells[:, :2].amin(dim=0).cpu().numpy()
        hi = cells[:, :2].amax(dim=0).cpu().numpy()
        self._bin_geometry(lo[0], hi[0], lo[1], hi[1])
        nb = self.nbx * self.nby
        counts = torch.zeros(nb, dtype=torch.int32, device=dev)
        args = (cells.data_ptr(), self.n_cells, self.bin_x0, self.bin_y0, self.bin_size, self.nbx, self.nby)
        _lib.check(self.lib.mlb_cells_bin(*args, 0, counts.data_ptr(), None, None, None, None, _stream_ptr()), "mlb_cells_bin")
        bin_start = torch.zeros(nb + 1, dtype=torch.int32, device=dev)
        bin_start[1:] = torch.cumsum(counts, dim=0, dtype=torch.int32)
        cursor = bin_start[:-1].clone()
        cx = torch.empty(self.n_cells, dtype=torch.float64, device=dev)
        cy = torch.empty(self.n_cells, dtype=torch.float64, device=dev)
        which = torch.empty(self.n_cells, dtype=torch.int32, device=dev)
        orig = torch.empty(self.n_cells, dtype=torch.int32, device=dev)
        _lib.check(self.lib.mlb_cells_bin(*args, 1, cursor.data_ptr(), cx.data_ptr(), cy.data_ptr(), which.data_ptr(),
                                          orig.data_ptr(), _stream_ptr()), "mlb_cells_bin")
        self._keep.update(cell_x=cx, cell_y=cy, cell_which=which, cell_orig=orig, bin_start=bin_start)

    def _tie_tables(self):
        """Host copies needed only when exact nearest-cell ties are resolved the reference's way."""
        if getattr(self, "_cells_host", None) is None:
            self._cells_host = self._cells_dev[:, 0:2].cpu().numpy().copy()
            orig = self._keep['cell_orig'].cpu().numpy().astype(np.int64)
            self._sorted_pos = np.empty(self.n_cells, dtype=np.int64)
            self._sorted_pos[orig] = np.arange(self.n_cells)

    def _pack_struct(self, pack, dev_arrays):
        s = _PackC()
        s.axes, s.values, s.values_f32, s.orders, s.order_map = (t.data_ptr() for t in dev_arrays)
        s.order_radius = pack.order_radius
        s.n_ux, s.n_uy, s.n_g = pack.n
        s.n_orders = len(pack.orders)
        for k in range(6):
            s.bounds[k] = pack.bounds[k]
        s.stats_slot = pack.stats_slot
        s.uniform01 = 1 if pack.uniform01 else 0
        s.u0_first, s.u1_first = pack.u_first
        s.u0_inv_step, s.u1_inv_step = pack.u_inv_step
        return s

    def _desc(self, d_x, d_y, source_x, source_y, source_z, source_pol, dipole_moment):
        L = _LensC()
        L.x_pts, L.y_pts, L.nx, L.ny = d_x.data_ptr(), d_y.data_ptr(), d_x.numel(), d_y.numel()
        k = self._keep
        L.ring_boundary, L.r_center = k['ring_boundary'].data_ptr(), k['r_center'].data_ptr()
        L.grating_period, L.num_around = k['grating_period'].data_ptr(), k['num_around'].data_ptr()
        L.gc_index = k['gc_index'].data_ptr()
        L.n_rings, L.n_packs = self.n_rings, len(self.packs)
        for g, p in enumerate(self.packs):
            L.packs[g] = self._pack_struct(p, self._pack_dev[g])
        L.n_cells, L.nbx, L.nby = self.n_cells, self.nbx, self.nby
        L.bin_x0, L.bin_y0, L.bin_size = self.bin_x0, self.bin_y0, self.bin_size
        if self.n_cells:
            L.cell_x, L.cell_y = k['cell_x'].data_ptr(), k['cell_y'].data_ptr()
            L.cell_which, L.cell_orig = k['cell_which'].data_ptr(), k['cell_orig'].data_ptr()
            L.bin_start = k['bin_start'].data_ptr()
        L.hex = self._pack_struct(self.hex_pack, self._pack_dev[-1])
        L.hex_x_period, L.hex_y_period = self.hex_x_period, self.hex_y_period
        L.plane_wave = 1 if source_z == -inf else 0
        L.source_x, L.source_y = float(source_x), float(source_y)
        L.source_z = -1.0 if source_z == -inf else float(source_z)
        L.source_pol = {'x': 0, 'y': 1, 'z': 2}[source_pol]
        L.wavelength, L.n_glass, L.dipole_moment = self.wavelength, float(self.n_glass), float(dipole_moment)
        L.c0, L.Z0 = c0, Z0
        L.ring_aux, L.ring_aux_f32 = k['ring_aux'].data_ptr(), k['ring_aux_f32'].data_ptr()
        L.ring_lut, L.n_lut, L.lut_r_max = k['ring_lut'].data_ptr(), self.n_lut, self.lens_max_r
        if k.get('ring_tables') is not None:
            L.ring_tables, L.ring_table_stride = k['ring_tables'].data_ptr(), self.ring_table_stride
        return L

    def default_grid(self):
        """Sample grid of nearfield.py:95-104: spacing ~ lambda/2.2, FFT-friendly count."""
        n = good_fft_number(2 * self.lens_max_r / (self.wavelength / 2.2))
        return np.linspace(-self.lens_max_r, self.lens_max_r, num=n)

    def _coords(self, pts):
        """Device copy of a coordinate list, cached on its bytes (a sweep calls run() with the same grid)."""
        a = np.ascontiguousarray(pts, dtype=np.float64)
        key = hash(a.tobytes())
        cache = self.__dict__.setdefault("_coord_cache", {})
        hit = cache.get(key)
        if hit is None or hit.numel() != a.size:
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            hit = cache[key] = torch.from_numpy(a).to(self.device)
        return hit

    def run(self, source_x, source_y, source_z, source_pol, x_pts, y_pts, dipole_moment=1e-30,
            out_dtype=torch.complex64, verbose=False, out=None, dxdy=None, check=True, ties="fast"):
        """Launch the fused assembly kernel.  Returns (fields (4, nx, ld) device tensor -- logical
        view [..., :ny] --, power device scalar float64).  Raises the reference's ValueErrors.
        x_pts / y_pts need not be uniform here (a multi-GPU rank passes ITS rows of the grid, slab.py); then
        `dxdy` gives the area element of the incident-power sum (:476-477) explicitly.
        check=False does not read the bounds-violation flag back (no host synchronisation: the launch stays
        asynchronous, e.g. inside a pipelined step); call check_violation() later -- it raises the same ValueError.
        ties: a centre sample EXACTLY equidistant from two hex cells (they sit on symmetry lines of symmetric grids) gets
        the highest cell row with "fast"; with "reference" those samples -- the kernel reports them -- are re-assembled
        with the cell the reference's own scipy cKDTree.query call returns (nearfield.py:363-364: whichever its
        traversal meets first), which makes the result equal to the reference's there as well.  Synchronises."""
        dev = self.device
        d_x, d_y = self._coords(x_pts), self._coords(y_pts)
        nx, ny = d_x.numel(), d_y.numel()
        ld = ny + (ny & 1)
        if out is None:
            out = torch.zeros((4, nx, ld), dtype=out_dtype, device=dev)
        assert out.dtype in (torch.complex64, torch.complex128) and tuple(out.shape) == (4, nx, ld)
        nblocks = self.lib.mlb_nearfield_blocks(nx, ny)
        block_sums = torch.empty(nblocks, dtype=torch.float64, device=dev)
        power = torch.empty(1, dtype=torch.float64, device=dev)
        if check:
            violation = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            if getattr(self, "_deferred", None) is None:
                self._deferred = torch.zeros(1, dtype=torch.int32, device=dev)
            violation = self._deferred
            self._deferred_args = (source_x, source_y, source_z, source_pol, x_pts, y_pts, dipole_moment)
        L = self._desc(d_x, d_y, source_x, source_y, source_z, source_pol, dipole_moment)
        stats = None

        def launch(want_stats):
            nonlocal stats
            sptr = None
            if want_stats:
                init = np.zeros((max(self.n_stats, 1), STATS_PER_ORDER), dtype=np.int64)
                init[:, 1::2] = _enc_inf(+1)
                init[:, 2::2] = _enc_inf(-1)
                init[:, 7] = 0
                stats = torch.from_numpy(init).to(dev)
                sptr = stats.data_ptr()
            if tie_buf is not None:
                tie_buf[0].zero_()
                rc = self.lib.mlb_nearfield_assemble_ties(
                    C.byref(L), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), ld,
                    1 if out.dtype == torch.complex128 else 0, block_sums.data_ptr(), sptr, 1 if want_stats else 0,
                    violation.data_ptr(), tie_buf[0].data_ptr(), tie_buf[1].data_ptr(), tie_buf[1].numel(), _stream_ptr())
                _lib.check(rc, "mlb_nearfield_assemble_ties")
                return
            rc = self.lib.mlb_nearfield_assemble(C.byref(L), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                                 out[3].data_ptr(), ld, 1 if out.dtype == torch.complex128 else 0,
                                                 block_sums.data_ptr(), sptr, 1 if want_stats else 0,
                                                 violation.data_ptr(), _stream_ptr())
            _lib.check(rc, "mlb_nearfield_assemble")

        assert ties in ("fast", "reference")
        tie_buf = None
        if ties == "reference":
            cap = 1 << 20
            tie_buf = (torch.zeros(1, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev))
        launch(bool(verbose))
        if check and (int(violation.item()) != 0 or verbose):
            if stats is None:
                violation.zero_()
                launch(True)                 # slow path: collect min/max for the reference's messages
            self._report(stats.cpu().numpy(), verbose)
        if tie_buf is not None:
            self._resolve_ties(L, out, ld, tie_buf, x_pts, y_pts, violation)
        if dxdy is None:
            dx = float(x_pts[1] - x_pts[0]) if nx > 1 else float('nan')
            dy = float(y_pts[1] - y_pts[0]) if ny > 1 else float('nan')
            dxdy = dx * dy
        rc = self.lib.mlb_sum_f64(block_sums.data_ptr(), nblocks, dxdy, power.data_ptr(), _stream_ptr())      # :476-477
        _lib.check(rc, "mlb_sum_f64")
        return out, power

    def _resolve_ties(self, L, out, ld, tie_buf, x_pts, y_pts, violation):
        """Re-assemble the samples whose index choice hinges on the last bit of a library call with the reference's own
        choice: centre samples exactly equidistant from two cells through the same
        ``cKDTree(lens_center_summary[:, 0:2]).query(points)`` call as nearfield.py:363-364, periphery samples on the
        boundary between two grating copies through numpy's ``round(arctan2(y, x) / angle_per_grating)`` (:119, :167-169)."""
        n = int(tie_buf[0].item())
        self.last_tie_count = n
        if n == 0:
            return
        if n > tie_buf[1].numel():
            raise _lib.MetalensB200Error("more index ties (%d) than the report list holds" % n)
        lin = np.sort(tie_buf[1][:n].cpu().numpy().astype(np.int64))
        ny = len(y_pts)
        px = np.asarray(x_pts, dtype=np.float64)[lin // ny]
        py = np.asarray(y_pts, dtype=np.float64)[lin % ny]
        k = self._keep
        bounds = k['ring_boundary'].cpu().numpy()
        ring = np.searchsorted(bounds, np.sqrt(px ** 2 + py ** 2)) - 1                           # :118, :125-126
        forced = np.zeros(n, dtype=np.int64)
        in_center = ring == -1
        self.last_tie_classes = (int(in_center.sum()), int((~in_center).sum()))
        if in_center.any():
            from scipy.spatial import cKDTree            # the reference's own dependency (nearfield.py:17)
            self._tie_tables()
            if getattr(self, "_tree", None) is None:
                self._tree = cKDTree(self._cells_host)
            winner = self._tree.query(np.stack((px[in_center], py[in_center]), axis=1))[1]      # original row numbers
            forced[in_center] = self._sorted_pos[winner]
        on_ring = ~in_center
        if on_ring.any():
            num = k['num_around'].cpu().numpy()[np.clip(ring[on_ring], 0, self.n_rings - 1)]
            apg = 2 * np.pi / num                                                                 # :161
            forced[on_ring] = np.round(np.arctan2(py[on_ring], px[on_ring]) / apg).astype(np.int64)   # :119, :167
        d_lin = torch.from_numpy(lin.astype(np.int32)).to(self.device)
        d_forced = torch.from_numpy(forced.astype(np.int32)).to(self.device)
        rc = self.lib.mlb_nearfield_fixup(C.byref(L), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                          out[3].data_ptr(), ld, 1 if out.dtype == torch.complex128 else 0,
                                          d_lin.data_ptr(), d_forced.data_ptr(), n, violation.data_ptr(), _stream_ptr())
        _lib.check(rc, "mlb_nearfield_fixup")

    def check_violation(self):
        """Deferred bounds check of run(check=False) calls: raises the reference's ValueError (:294-305, :412-419)
        if any of them left the tables (by re-running the last one with statistics)."""
        flag = getattr(self, "_deferred", None)
        if flag is not None and int(flag.item()) != 0:
            flag.zero_()
            a = self._deferred_args
            self.run(a[0], a[1], a[2], a[3], a[4], a[5], dipole_moment=a[6], check=True)
            raise _lib.MetalensB200Error("an earlier run(check=False) call left the interpolation tables")

    def _report(self, stats, verbose):
        """Replay the reference's per-order bounds checks in its own loop order
        (nearfield.py:294-305 for each collection, then :412-419 for the centre)."""
        for gi, p in enumerate(self.packs):
            for o, (ox, oy) in enumerate(p.orders):
                row = stats[p.stats_slot + o]
                if row[0] == 0:
                    continue
                if verbose:
                    print('diffraction order', (ox, oy), 'of gc', gi, '; applies at', int(row[0]), 'points', flush=True)
                v = _dec(row[1:7])
                b = p.bounds
                if v[0] < b[0]:
                    raise ValueError('need to calculate at smaller ux!', v[0], b[0])
                if v[1] > b[1]:
                    raise ValueError('need to calculate at bigger ux!', v[1], b[1])
                if v[2] < b[2]:
                    raise ValueError('need to calculate at smaller uy!', v[2], b[2])
                if v[3] > b[3]:
                    raise ValueError('need to calculate at bigger uy!', v[3], b[3])
                if v[4] < b[4]:
                    raise ValueError('need to calculate at smaller grating_period!', v[4] / nm, b[4] / nm)
                if v[5] > b[5]:
                    raise ValueError('need to calculate at bigger grating_period!', v[5] / nm, b[5] / nm)
        p = self.hex_pack
        for o, (ox, oy) in enumerate(p.orders):
            row = stats[p.stats_slot + o]
            if row[0] == 0:
                continue
            if verbose:
                print('diffraction order', (ox, oy), 'of center; applies at', int(row[0]), 'points', flush=True)
            v = _dec(row[1:7])
            b = p.bounds
            if v[0] < b[0]:
                raise ValueError('need to calculate at smaller ux!', v[0], b[0])
            if v[1] > b[1]:
                raise ValueError('need to calculate at bigger ux!', v[1], b[1])
            if v[2] < b[2]:
                raise ValueError('need to calculate at smaller uy!', v[2], b[2])
            if v[3] > b[3]:
                raise ValueError('need to calculate at bigger uy!', v[3], b[3])


_PLAN_CACHE = {}


def _plan_for(wavelength, periphery, center, hexgridset):
    # the exact wavelength (k_vac, k_glass, the default grid depend on it; only the TABLE lookup uses the rounded nm,
    # nearfield.py:86) and the device the plan's buffers live on
    if not torch.cuda.is_available():
        raise _lib.MetalensB200Error("metalens_b200 needs a CUDA device (no CPU fallback)")
    key = (float(wavelength), torch.cuda.current_device(), id(periphery), id(center), id(hexgridset),
           id(getattr(hexgridset, "interpolators", None)),
           tuple(id(getattr(gc, "interpolators", None)) for gc in periphery['gratingcollection_list']))
    hit = _PLAN_CACHE.get(key)
    if hit is None or hit[1] is not periphery or hit[2] is not center or hit[3] is not hexgridset:
        if len(_PLAN_CACHE) >= 4:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
        hit = (NearfieldPlan(wavelength, periphery, center, hexgridset), periphery, center, hexgridset)
        _PLAN_CACHE[key] = hit
    return hit[0]


def build_nearfield(source_x, source_y, source_z, source_pol, wavelength,
                    lens_periphery_summary, lens_center_summary, hexgridset,
                    x_pts=None, y_pts=None, dipole_moment=1e-30, verbose=False):
    """Drop-in for the reference's ``nearfield.build_nearfield`` (nearfield.py:66-480).

    Same arguments (``dipole_moment`` in SI: C*m for a point source, V/m for the plane wave
    selected by ``source_z = -inf``), same return tuple
    ``(Ex, Ey, Hx, Hy, x_pts, y_pts, power_passing_through_lens, n_glass)`` with complex128
    (nx, ny) arrays, same AssertionErrors (:84-85, :106-109, :224) and the same
    ``ValueError('need to calculate at smaller ux!', value, bound)`` family (:294-305,
    :412-419).  ``verbose=True`` reproduces the reference's per-order progress prints.
    """
    assert source_z < 0                                                        # :84
    assert source_pol in ('x', 'y', 'z')                                       # :85
    def check_axis(l):                                                         # :106-109
        steps = np.diff(np.asarray(l, dtype=float))
        assert 0 < steps[0] < wavelength / 2
        assert steps.max() - steps.min() <= 1e-9 * np.abs(steps).max()
    for l in (x_pts, y_pts):               # the caller's axes are judged before any device work, like the reference
        if l is not None:
            check_axis(l)
    if source_z == -inf:
        assert source_pol != 'z'                                               # :224
    plan = _plan_for(wavelength, lens_periphery_summary, lens_center_summary, hexgridset)
    if x_pts is None:
        x_pts = plan.default_grid()                                            # :95-99
        check_axis(x_pts)
    if y_pts is None:
        y_pts = plan.default_grid()                                            # :100-104
        check_axis(y_pts)
    out, power = plan.run(source_x, source_y, source_z, source_pol, x_pts, y_pts, dipole_moment=dipole_moment,
                          out_dtype=torch.complex128, verbose=verbose, ties="reference")
    ny = len(y_pts)
    host = out[:, :, :ny].cpu().numpy()
    return host[0], host[1], host[2], host[3], x_pts, y_pts, float(power.item()), plan.n_glass


def build_nearfield_big(source_x, source_y, source_z, source_pol, wavelength,
                        lens_periphery_summary, lens_center_summary, hexgridset,
                        x_pts=None, y_pts=None, dipole_moment=1e-30, verbose=False):
    """Drop-in for ``nearfield.build_nearfield_big`` (nearfield.py:482-516).  The reference
    slices y into slabs of 1e7/len(x_pts) columns only to bound host RAM; here every slab is one
    kernel launch into the same arrays, so results equal the single-call ones.  As in the
    reference, ``x_pts`` and ``y_pts`` are required (it reads ``x_pts.size`` at :489)."""
    x_pts = np.asarray(x_pts)
    y_pts = np.asarray(y_pts)
    per = max(2, int(1e7 / x_pts.size))                                       # :488-489
    parts, power, n_glass = [], 0, None
    for start in range(0, y_pts.size, per):
        if verbose:
            print('running y-index', start, 'to', min(start + per, y_pts.size), 'out of', y_pts.size, flush=True)
        res = build_nearfield(source_x, source_y, source_z, source_pol, wavelength, lens_periphery_summary,
                              lens_center_summary, hexgridset, x_pts=x_pts, y_pts=y_pts[start:start + per],
                              dipole_moment=dipole_moment, verbose=verbose)
        parts.append(res[:4])
        power += res[6]
        n_glass = res[7]
    Ex, Ey, Hx, Hy = (np.concatenate([p[i] for p in parts], axis=1) for i in range(4))
    return Ex, Ey, Hx, Hy, x_pts, y_pts, power, n_glass
