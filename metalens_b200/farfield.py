"""Near-field -> far-field transform on the B200 (hot path A of SURVEY.md section 8).

Python host side of the reference's ``nearfield_farfield.py``.  Three entry points:

* :func:`farfield_from_nearfield` -- the reference's own signature and return tuple
  (``nearfield_farfield.py:14-75``): takes ``fft2(fftshift(field))`` arrays, so only
  the radiated-power epilogue (``:77-191``) runs on the GPU.  Strict drop-in.
* :func:`farfield_from_fields` -- takes the real-space aperture fields and does the
  aperture sum on the GPU as well (replaces the caller-side FFTs of ``:18-20``),
  on the reference's FFT-bin grid, every ``stride``-th bin of it, or an arbitrary
  direction-cosine grid.
* :class:`FarfieldPlan` -- the reusable device-resident engine behind both: twiddle
  tables, workspaces and kernel launches through the C-ABI (``include/metalens_b200.h``).

torch is used only for device buffers, streams and copies; every arithmetic step
is a kernel of ``libmetalens_b200.so``.  There is no CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib, hostmem
from .units import Z0


TWO_PASS_TRANSPOSE = False      # experiment switch, see FarfieldPlan._build


def fft_bin_direction_cosines(num, spacing, wavelength, n_glass):
    """Un-shifted direction cosines of the FFT bins, same arithmetic as the reference
    (nearfield_farfield.py:35-39) so the NaN mask of evanescent bins is bit-identical."""
    lam = wavelength / n_glass
    u = np.arange(num) * lam / (spacing * num)
    u[u > u.max() / 2] -= lam / spacing
    return u


def _check_axis(pts, wavelength):
    """Grid validation, nearfield_farfield.py:26-30 (AssertionError like the reference)."""
    pts = np.asarray(pts, dtype=float)
    steps = np.diff(pts)
    assert 0 < steps[0] < wavelength / 2
    assert steps.max() - steps.min() <= 1e-9 * np.abs(steps).max()


def _even(n):
    return n + (n & 1)


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _c64_buffer(rows, cols, device):
    """Zeroed complex64 [rows][ld] device buffer with an even pitch; returns the
    full (pitched) tensor -- its logical view is buf[:, :cols]."""
    return torch.zeros((rows, _even(cols)), dtype=torch.complex64, device=device)


class FarfieldPlan:
    """Device-resident NF->FF engine for one aperture geometry and one far-field grid.

    Parameters
    ----------
    shape : (Mx, My) aperture samples (axis 0 = x, axis 1 = y, C order; nearfield.py:117)
    dxp, dyp : sample spacings
    wavelength, n_glass : vacuum wavelength and substrate index (direction cosines are
        *in glass*, nearfield_farfield.py:33-36)
    stride : int or (sx, sy) -- far-field grid = every stride-th bin of the reference's
        fftshifted FFT-bin grid (stride 1 = the reference grid itself)
    ux, uy : explicit direction-cosine lists (arbitrary grid); excludes `stride`
    method : 'auto' | 'dense' | 'fold' | 'fft' | 'czt' | 'tc'
        dense -- two-stage separable tiled complex reduction over the full aperture
        czt   -- chirp-z (Bluestein) aperture sum for UNIFORM ux / uy lists that are not FFT bins (a "zoomed" far field):
                 the sum along each axis becomes a convolution with a chirp, run on the FFT passes at a power-of-two
                 length L >= M + K - 1 <= 8192 (csrc/czt.cu); float32-exact like 'dense' at a small fraction of its work
        fold  -- exact aperture fold (HBM-bound) followed by the dense reduction on the
                 folded (Mx/sx x My/sy) aperture; needs an FFT-bin-stride grid with
                 sx | Mx//2 and sy | My//2
        fft   -- shared-memory row/column FFT passes, the row pass folding the aperture while it
                 loads (stride > 1): the reference's own algorithm, every kernel memory-bound,
                 the aperture read from HBM exactly once; folded sizes of the form 2^a 3^b 5^c
                 (what good_fft_number() yields) up to 8192 -- powers of two take the tuned
                 kernels (TMA-fed row pass), other sizes the mixed-radix ones
        tc    -- the dense separable reduction on the tensor cores: tcgen05 UMMA (TF32 operands split
                 hi/lo, three products per term -> fp32-class accuracy), TMEM accumulators, TMA-fed
        auto  -- fft if eligible, else fold, else (explicit uniform ux / uy) czt, else dense
    p_dtype : torch.float32 (north-star output type) or torch.float64
    rows : optional (row0, row1): compute only that slab of far-field rows (ux indices) -- the
        multi-GPU tile of metalens_b200/sharding.py.  Supported by 'dense' and 'fold', whose work
        scales with the slab; the FFT passes produce all rows at once.
    fuse_power : with method 'fft', float32 P and power-of-two column lengths 256..1024, run() uses the
        fused column-pass + power kernel (mlb_fft_cols_power): the 4 x Kx x Ky aperture sums then stay in
        registers; amplitudes() re-runs the unfused passes when asked.  False = always separate kernels,
        'always' = fused for every length the kernel supports (up to 8192).
    """

    def __init__(self, shape, dxp, dyp, wavelength, n_glass, stride=None, ux=None, uy=None,
                 method="auto", p_dtype=torch.float32, device=None, rows=None, fuse_power=True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MetalensB200Error("metalens_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.Mx, self.My = int(shape[0]), int(shape[1])
        self.dxp, self.dyp = float(dxp), float(dyp)
        self.wavelength, self.n_glass = float(wavelength), float(n_glass)
        assert p_dtype in (torch.float32, torch.float64)
        self.p_dtype = p_dtype
        self._want_fused = fuse_power if fuse_power == "always" else bool(fuse_power)

        if ux is None and uy is None:
            stride = 1 if stride is None else stride
            sx, sy = (stride, stride) if np.isscalar(stride) else stride
            self.sx, self.sy = int(sx), int(sy)
            ux = np.fft.fftshift(fft_bin_direction_cosines(self.Mx, self.dxp, wavelength, n_glass))[::self.sx]
            uy = np.fft.fftshift(fft_bin_direction_cosines(self.My, self.dyp, wavelength, n_glass))[::self.sy]
            self.fft_bin_grid = True
        else:
            assert stride is None and ux is not None and uy is not None
            self.sx = self.sy = None
            self.fft_bin_grid = False
        self.ux = np.ascontiguousarray(ux, dtype=np.float64)
        self.uy = np.ascontiguousarray(uy, dtype=np.float64)
        self.Kx_full = self.ux.size
        self.rows = None if rows is None else (int(rows[0]), int(rows[1]))
        if self.rows is not None:
            assert 0 <= self.rows[0] < self.rows[1] <= self.Kx_full
            self.ux = np.ascontiguousarray(self.ux[self.rows[0]:self.rows[1]])
        self.Kx, self.Ky = self.ux.size, self.uy.size

        can_fold = (self.fft_bin_grid and self.Mx % self.sx == 0 and self.My % self.sy == 0
                    and (self.Mx // 2) % self.sx == 0 and (self.My // 2) % self.sy == 0
                    and (self.sx > 1 or self.sy > 1))
        def smooth(n):                 # 2^a 3^b 5^c: what good_fft_number() produces (nearfield.py:30-36)
            if n < 2:
                return False
            for f in (2, 3, 5):
                while n % f == 0:
                    n //= f
            return n == 1
        can_fft = False
        if self.fft_bin_grid and self.Mx % self.sx == 0 and self.My % self.sy == 0:
            k1, k2 = self.Mx // self.sx, self.My // self.sy
            nmax = self.lib.mlb_fft_max_length()
            can_fft = (smooth(k1) and smooth(k2) and k1 <= nmax and k2 <= nmax
                       and (can_fold or (self.sx == 1 and self.sy == 1)))
        if self.rows is not None and self.rows != (0, self.Kx_full):
            if method == "fft":
                raise ValueError("row slabs need method 'dense' or 'fold' (the FFT passes yield all rows)")
            can_fft = False
        def uniform(u):
            d = np.diff(u)
            return bool(u.size >= 2 and d[0] != 0 and np.all(np.abs(d - d[0]) <= 1e-9 * abs(d[0])))

        def czt_len(m, k):
            n = 1
            while n < m + k - 1:
                n *= 2
            return n
        self.czt_L = (czt_len(self.Mx, self.Kx), czt_len(self.My, self.Ky))
        nmax_czt = self.lib.mlb_fft_max_length()
        can_czt = (uniform(self.ux) and uniform(self.uy) and min(self.czt_L) >= 16 and max(self.czt_L) <= nmax_czt
                   and max(self.czt_L) <= 65535)
        if method == "auto":
            method = "fft" if can_fft else ("fold" if can_fold else ("czt" if (can_czt and not self.fft_bin_grid) else "dense"))
        if method == "czt" and not can_czt:
            raise ValueError("czt needs uniformly spaced ux and uy lists and M + K - 1 <= %d per axis" % nmax_czt)
        if method == "fold" and not can_fold:
            raise ValueError("fold needs an FFT-bin-stride grid with stride dividing M and M//2")
        if method == "fft" and not can_fft:
            raise ValueError("fft needs an FFT-bin(-stride) grid whose folded sizes are 2^a 3^b 5^c and <= %d"
                             % self.lib.mlb_fft_max_length())
        assert method in ("dense", "fold", "fft", "tc", "czt")
        self.method = method
        self._build()

    # ------------------------------------------------------------------ setup
    def _twiddle(self, coord, u, scale):
        """[len(coord)][even(len(u))] complex64 table exp(i*pi*scale*coord*u)."""
        dev = self.device
        c = torch.from_numpy(np.ascontiguousarray(coord, dtype=np.float64)).to(dev)
        v = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(dev)
        out = _c64_buffer(c.numel(), v.numel(), dev)
        rc = self.lib.mlb_twiddle_build(c.data_ptr(), c.numel(), v.data_ptr(), v.numel(), float(scale),
                                        out.data_ptr(), out.shape[1], _stream_ptr())
        _lib.check(rc, "mlb_twiddle_build")
        return out

    def _build(self):
        dev = self.device
        Mx, My, Kx, Ky = self.Mx, self.My, self.Kx, self.Ky
        if self.method == "dense":
            # phase origin = the sample fftshift() moves to index 0: M - M//2
            ox, oy = Mx - Mx // 2, My - My // 2
            scale = -2.0 * self.n_glass / self.wavelength
            self.AxT = self._twiddle((np.arange(Mx) - ox) * self.dxp, self.ux, scale)   # [Mx][Kx]
            self.Ay = self._twiddle((np.arange(My) - oy) * self.dyp, self.uy, scale)    # [My][Ky]
            self.Rx, self.Ry = Mx, My            # size of the aperture the reduction runs over
            self.G = None
        elif self.method == "tc":
            self._build_tc()
        elif self.method == "czt":
            self._build_czt()
        elif self.method == "fft":
            K1, K2 = Mx // self.sx, My // self.sy
            assert K1 == Kx and K2 == Ky
            self.Rx, self.Ry = K1, K2
            self.G = None          # the row pass folds while loading: the folded aperture never exists in memory
            # Row pass followed by a strided column pass.  (Two transposing row passes, both streaming
            # contiguous rows through TMA, are implemented -- transpose_out of mlb_fft_rows -- but measured
            # slower on B200: 98 + 54 us instead of 92 + 45 us at 1024^2, the 8-byte transposed stores and the
            # one-row-at-a-time FFT latency of the second pass cost more than the strided reads they avoid.)
            self.two_pass_t = bool(TWO_PASS_TRANSPOSE and self.lib.mlb_fft_rows_can_transpose(K1)
                                   and self.lib.mlb_fft_rows_can_transpose(K2))
            self.W = [_c64_buffer(K2, K1, dev) if self.two_pass_t else _c64_buffer(K1, K2, dev) for _ in range(4)]
            self.tw1 = torch.empty(2 * K1, dtype=torch.complex64, device=dev)     # plain + staged tables
            self.tw2 = torch.empty(2 * K2, dtype=torch.complex64, device=dev)
            for t, n in ((self.tw1, K1), (self.tw2, K2)):
                _lib.check(self.lib.mlb_fft_twiddle(n, t.data_ptr(), _stream_ptr()), "mlb_fft_twiddle")
            self.AxT = self.Ay = None
            self.done_counter = torch.zeros(2, dtype=torch.int32, device=dev)     # finished-CTA counter of the fused column+power pass
            # fused column pass + power: measured faster than column pass + epilogue up to 1024-point columns
            # (above that the fused kernel's register footprint costs more than the saved round trip)
            self.fused = bool(self._want_fused and not self.two_pass_t and self.p_dtype == torch.float32
                              and self.lib.mlb_fft_cols_power_blocks(K1, K2) > 0
                              and (K1 <= 1024 or self._want_fused == "always"))
        else:
            # folded aperture (K1 x K2) and exact integer DFT twiddles exp(-2 pi i p q / K)
            K1, K2 = Mx // self.sx, My // self.sy
            assert K1 == self.Kx_full and K2 == Ky
            qx = (np.arange(K1) - (Mx // 2) // self.sx) % K1     # un-shifted bin number of output q'
            qy = (np.arange(K2) - (My // 2) // self.sy) % K2
            if self.rows is not None:
                qx = qx[self.rows[0]:self.rows[1]]
            self.AxT = self._twiddle(np.arange(K1), qx, -2.0 / K1)
            self.Ay = self._twiddle(np.arange(K2), qy, -2.0 / K2)
            self.Rx, self.Ry = K1, K2
            self.G = [_c64_buffer(K1, K2, dev) for _ in range(4)]
        self.UT = ([_c64_buffer(self.Ry, Kx, dev) for _ in range(4)]     # stage-1 output, [m2][i]
                   if self.method in ("dense", "fold") else None)
        self.fused = bool(getattr(self, "fused", False))
        self._Fhat = None if self.fused else [_c64_buffer(Kx, Ky, dev) for _ in range(4)]   # aperture sums, [i][j]
        self._fields_last = None
        self._steps_cache = {}
        self._staged = False
        self.P = torch.empty((Kx, Ky), dtype=self.p_dtype, device=dev)
        self.nblocks = self.lib.mlb_ff_epilogue_blocks(Kx, Ky)
        self.block_sums = torch.empty(max(self.nblocks, 2 * Ky + 16), dtype=torch.float64, device=dev)
        self.total = torch.zeros(1, dtype=torch.float64, device=dev)
        self.d_ux = torch.from_numpy(self.ux).to(dev)
        self.d_uy = torch.from_numpy(self.uy).to(dev)
        self.dux = float(self.ux[1] - self.ux[0]) if Kx > 1 else float("nan")
        self.duy = float(self.uy[1] - self.uy[0]) if Ky > 1 else float("nan")
        self._staging = None
        self._pinned = None

    def _build_czt(self):
        """Chirp tables, their transforms and the work buffers of method 'czt' (csrc/czt.cu)."""
        dev, lib = self.device, self.lib
        self.Rx, self.Ry = self.Mx, self.My
        self.G = None
        self.AxT = self.Ay = None
        self.czt = {}
        for axis, (M, K, L, d, u) in {"x": (self.Mx, self.Kx, self.czt_L[0], self.dxp, self.ux),
                                      "y": (self.My, self.Ky, self.czt_L[1], self.dyp, self.uy)}.items():
            o = M - M // 2                                    # the sample fftshift() moves to index 0
            du = float(u[1] - u[0])
            t_lin = 2.0 * self.n_glass * d * float(u[0]) / self.wavelength
            t_quad = 2.0 * self.n_glass * d * du / self.wavelength
            pre = torch.empty(M + (M & 1), dtype=torch.complex64, device=dev)
            kern = torch.empty((1, L), dtype=torch.complex64, device=dev)
            post = torch.empty(K + (K & 1), dtype=torch.complex64, device=dev)
            _lib.check(lib.mlb_czt_chirps(M, o, K, L, t_lin, t_quad, pre.data_ptr(), kern.data_ptr(), post.data_ptr(),
                                          _stream_ptr()), "mlb_czt_chirps")
            tw = torch.empty(2 * L, dtype=torch.complex64, device=dev)
            _lib.check(lib.mlb_fft_twiddle(L, tw.data_ptr(), _stream_ptr()), "mlb_fft_twiddle")
            khat = torch.empty((1, L), dtype=torch.complex64, device=dev)
            pk, k1 = _lib.ptr_array([kern])
            ph, k2 = _lib.ptr_array([khat])
            _lib.check(lib.mlb_fft_rows(pk, L, ph, L, 1, L, 1, 1, tw.data_ptr(), 0, 0, 0, 0, 1, _stream_ptr()), "mlb_fft_rows(chirp)")
            self.czt[axis] = dict(o=o, L=L, pre=pre, post=post, khat=khat, tw=tw)
        Lx, Ly = self.czt_L
        self.czt_R = [torch.empty((self.Mx, Ly), dtype=torch.complex64, device=dev) for _ in range(4)]
        kyp = _even(self.Ky)
        self.czt_C = [[torch.empty((Lx, kyp), dtype=torch.complex64, device=dev) for _ in range(4)] for _ in range(2)]

    def _steps_czt(self, ops, ld):
        lib = self.lib
        Mx, My, Kx, Ky = self.Mx, self.My, self.Kx, self.Ky
        Lx, Ly = self.czt_L
        X, Y = self.czt["x"], self.czt["y"]
        pin, k0 = _lib.ptr_array(ops)
        pR, k1 = _lib.ptr_array(self.czt_R)
        pC0, k2 = _lib.ptr_array(self.czt_C[0])
        pC1, k3 = _lib.ptr_array(self.czt_C[1])
        pF, k4 = _lib.ptr_array(self.Fhat)
        ldc, ldf = self.czt_C[0][0].shape[1], self.Fhat[0].shape[1]

        def pw(i, ldi, o, ldo, rows_out, cols_out, rows_valid, cols_valid, roff, coff, rtab, ctab, cin, cout):
            _lib.check(lib.mlb_czt_pointwise(i, ldi, o, ldo, rows_out, cols_out, rows_valid, cols_valid, roff, coff,
                                             None if rtab is None else rtab.data_ptr(),
                                             None if ctab is None else ctab.data_ptr(), cin, cout, 4, _stream_ptr()),
                       "mlb_czt_pointwise")

        def rows(keep=(k0, k1)):            # along y: pad . pre -> FFT -> . FFT(chirp), conj -> FFT -> (crop in cols())
            pw(pin, ld, pR, Ly, Mx, Ly, Mx, My, 0, 0, None, Y["pre"], 0, 0)
            _lib.check(lib.mlb_fft_rows(pR, Ly, pR, Ly, Mx, Ly, 1, 1, Y["tw"].data_ptr(), 0, 0, 0, 0, 4, _stream_ptr()), "mlb_fft_rows(czt)")
            pw(pR, Ly, pR, Ly, Mx, Ly, Mx, Ly, 0, 0, None, Y["khat"], 0, 1)
            _lib.check(lib.mlb_fft_rows(pR, Ly, pR, Ly, Mx, Ly, 1, 1, Y["tw"].data_ptr(), 0, 0, 0, 0, 4, _stream_ptr()), "mlb_fft_rows(czt)")

        def cols(keep=(k2, k3, k4)):        # crop . post_y . pre_x, zero-padded to Lx rows -> the same along x -> crop . post_x
            pw(pR, Ly, pC0, ldc, Lx, Ky, Mx, Ky, 0, Y["o"], X["pre"], Y["post"], 1, 0)
            _lib.check(lib.mlb_fft_cols(pC0, ldc, pC1, ldc, Lx, Ky, X["tw"].data_ptr(), 0, 4, _stream_ptr()), "mlb_fft_cols(czt)")
            pw(pC1, ldc, pC1, ldc, Lx, Ky, Lx, Ky, 0, 0, X["khat"], None, 0, 1)
            _lib.check(lib.mlb_fft_cols(pC1, ldc, pC0, ldc, Lx, Ky, X["tw"].data_ptr(), 0, 4, _stream_ptr()), "mlb_fft_cols(czt)")
            pw(pC0, ldc, pF, ldf, Kx, Ky, Kx, Ky, X["o"], 0, X["post"], None, 1, 0)
        import math as _m
        return [("czt_rows", rows, 4 * 8 * (Mx * My + 5 * Mx * Ly), 4 * 2 * 5.0 * Mx * Ly * _m.log2(Ly)),
                ("czt_cols", cols, 4 * 8 * (Mx * Ly + 5 * Lx * Ky + Kx * Ky), 4 * 2 * 5.0 * Lx * Ky * _m.log2(Lx))]

    def _build_tc(self):
        """Operands of the tensor-core path (csrc/cgemm_tc.cu): y-first contraction
        T[m1][j] = sum_m2 J[m1][m2] Ay[m2][j];  F[i][j] = sum_m1 Ax[i][m1] T[m1][j]."""
        dev, lib = self.device, self.lib
        Mx, My, Kx, Ky = self.Mx, self.My, self.Kx, self.Ky
        self.Rx, self.Ry = Mx, My
        self.G = None
        ox, oy = Mx - Mx // 2, My - My // 2
        scale = -2.0 * self.n_glass / self.wavelength

        def pad4(n):
            return (n + 3) // 4 * 4

        def f32(rows, cols):
            return torch.zeros((rows, pad4(cols)), dtype=torch.float32, device=dev)

        def tw(coord, u, layout, rows):
            c = torch.from_numpy(np.ascontiguousarray(coord, dtype=np.float64)).to(dev)
            v = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(dev)
            hi, lo = f32(rows, 2 * c.numel()), f32(rows, 2 * c.numel())
            _lib.check(lib.mlb_twiddle_tf32(c.data_ptr(), c.numel(), v.data_ptr(), v.numel(), scale, layout,
                                            hi.data_ptr(), lo.data_ptr(), hi.shape[1], _stream_ptr()), "mlb_twiddle_tf32")
            return hi, lo
        # stage-1 B operand: embedding of Ay[m2][j] = e^{-ik y'_m2 uy_j}, rows 2j+q, columns 2m2+p
        self.AyB = tw((np.arange(My) - oy) * self.dyp, self.uy, 1, 2 * Ky)
        # stage-2 A operand: Ax[i][m1] = e^{-ik x'_m1 ux_i}, row i, columns (2m1, 2m1+1)
        self.AxA = tw((np.arange(Mx) - ox) * self.dxp, self.ux, 0, Kx)
        self.Jh = [f32(Mx, 2 * My) for _ in range(4)]
        self.Jl = [f32(Mx, 2 * My) for _ in range(4)]
        self.TBh = [f32(2 * Ky, 2 * Mx) for _ in range(4)]
        self.TBl = [f32(2 * Ky, 2 * Mx) for _ in range(4)]
        # partial sums of the chunked stage 1 (mlb_cgemm_tc_split): complex64 [Mx][Ky]
        self.tc_scratch = [_c64_buffer(Mx, Ky, dev) for _ in range(4)]
        self.AxT = self.Ay = None

    def _steps_tc(self, ops, ld):
        lib = self.lib
        Mx, My, Kx, Ky = self.Mx, self.My, self.Kx, self.Ky
        ldj = self.Jh[0].shape[1]

        def split():
            for f in range(4):
                _lib.check(lib.mlb_tf32_split(ops[f].data_ptr(), 2 * ld, self.Jh[f].data_ptr(), self.Jl[f].data_ptr(),
                                              ldj, Mx, 2 * My, _stream_ptr()), "mlb_tf32_split")

        pJh, k1 = _lib.ptr_array(self.Jh)
        pJl, k2 = _lib.ptr_array(self.Jl)
        pTh, k3 = _lib.ptr_array(self.TBh)
        pTl, k4 = _lib.ptr_array(self.TBl)
        pAyh, k5 = _lib.ptr_array([self.AyB[0]] * 4)
        pAyl, k6 = _lib.ptr_array([self.AyB[1]] * 4)
        pAxh, k7 = _lib.ptr_array([self.AxA[0]] * 4)
        pAxl, k8 = _lib.ptr_array([self.AxA[1]] * 4)
        pF, k9 = _lib.ptr_array(self.Fhat)
        pS, k10 = _lib.ptr_array(self.tc_scratch)

        # both stages with the contraction split into 512-deep chunks whose partial sums are added with round-to-nearest
        # (the tensor core's truncating accumulation is a bias that grows with the depth of a coherent sum)
        def stage1(keep=(k1, k2, k3, k4, k5, k6, k10)):  # all four fields in one launch per chunk
            _lib.check(lib.mlb_cgemm_tc_split(pJh, pJl, ldj, pAyh, pAyl, self.AyB[0].shape[1], Mx, Ky, My, 1,
                                              pTh, pTl, self.TBh[0].shape[1], 4, pS, self.tc_scratch[0].shape[1],
                                              _stream_ptr()), "mlb_cgemm_tc_split(stage 1)")

        def stage2(keep=(k7, k8, k9)):
            _lib.check(lib.mlb_cgemm_tc_split(pAxh, pAxl, self.AxA[0].shape[1], pTh, pTl, self.TBh[0].shape[1],
                                              Kx, Ky, Mx, 2, pF, None, self.Fhat[0].shape[1], 4, None, 0, _stream_ptr()),
                       "mlb_cgemm_tc_split(stage 2)")
        return [("tf32_split", split, 4 * 24 * Mx * My, 0.0),
                ("tc_stage1", stage1, 4 * (16 * Mx * My + 32 * Ky * Mx) + 16 * Ky * My, 32.0 * Mx * My * Ky),
                ("tc_stage2", stage2, 4 * (32 * Ky * Mx + 8 * Kx * Ky) + 16 * Kx * Mx, 32.0 * Kx * Mx * Ky)]

    @property
    def Fhat(self):
        """The four aperture-sum buffers [Kx][even(Ky)] (allocated on first use by fused plans)."""
        if self._Fhat is None:
            self._Fhat = [_c64_buffer(self.Kx, self.Ky, self.device) for _ in range(4)]
        return self._Fhat

    # ------------------------------------------------------------------ run
    def _as_operands(self, fields):
        """Return 4 pitched complex64 device tensors [Mx][even(My)] for the kernels;
        copies only when the caller's layout cannot be used in place."""
        out = []
        self._staged = False
        for idx, f in enumerate(fields):
            assert f.is_cuda and f.dtype == torch.complex64 and tuple(f.shape[-2:]) == (self.Mx, self.My), \
                "fields must be CUDA complex64 (Mx, My)"
            ok = (f.stride(-1) == 1 and f.stride(-2) % 2 == 0 and f.stride(-2) >= self.My
                  and f.data_ptr() % 16 == 0)
            if ok:
                out.append((f, f.stride(-2)))
            else:
                if self._staging is None:
                    self._staging = [_c64_buffer(self.Mx, self.My, self.device) for _ in range(4)]
                self._staging[idx][:, :self.My].copy_(f)
                self._staged = True
                out.append((self._staging[idx], self._staging[idx].shape[1]))
        lds = {ld for _, ld in out}
        assert len(lds) == 1, "the four fields must share one row pitch"
        return [t for t, _ in out], lds.pop()

    def steps(self, fields, fused=None, accumulate=False):
        """The kernel launches of one run() as (name, thunk, algorithmic bytes, algorithmic flops);
        run() executes them in order, bench.py times them one by one for the roofline.
        fused=False forces the separate column pass + epilogue (aperture sums stored in Fhat)."""
        fused = self.fused if fused is None else (bool(fused) and self.fused)
        # the launch thunks only depend on the operand addresses: cache them per set of field buffers
        key = (tuple((f.data_ptr(), f.stride(-2), f.stride(-1)) for f in fields), fused, bool(accumulate))
        hit = self._steps_cache.get(key)
        if hit is not None:
            return hit
        out = self._make_steps(fields, fused, accumulate)
        if not self._staged:                     # staged operands are copied on every call: never cached
            if len(self._steps_cache) >= 16:
                self._steps_cache.clear()
            self._steps_cache[key] = out
        return out

    def _make_steps(self, fields, fused, accumulate):
        lib = self.lib
        ops, ld = self._as_operands(fields)
        Kx, Ky, Rx, Ry = self.Kx, self.Ky, self.Rx, self.Ry
        out = []
        if self.method == "tc":
            return self._steps_tc(ops, ld) + [("epilogue", self.power, 36 * Kx * Ky, 0.0)]
        if self.method == "czt":
            return self._steps_czt(ops, ld) + [("epilogue", self.power, 36 * Kx * Ky, 0.0)]
        if self.method == "fold":
            pj, k1 = _lib.ptr_array(ops)
            pg, k2 = _lib.ptr_array(self.G)
            ldg = self.G[0].shape[1]

            def fold(pj=pj, pg=pg, ld=ld, keep=(k1, k2)):
                _lib.check(lib.mlb_fold(pj, ld, self.Mx, self.My, self.sx, self.sy, self.Mx // 2, self.My // 2,
                                        pg, ldg, 4, _stream_ptr()), "mlb_fold")
            out.append(("fold", fold, 32 * (self.Mx * self.My + Rx * Ry), 0.0))
            ops, ld, folded = self.G, ldg, True
        else:
            folded = False
        if self.method == "fft":
            # F[s q] = sum_p G[p] e^{-2 pi i q p / K},  G[p] = sum_t J[((p - M//2) mod K) + t K]; outputs are
            # stored at the fftshifted position q' = (q + (M//2)/s) mod K, the order of self.ux / self.uy
            h1, h2 = self.Mx // 2, self.My // 2
            roll_r, roll_c = h1 % Rx, h2 % Ry          # fftshift of the input (:18-20), modulo the fold
            assert not folded
            pi_, k3 = _lib.ptr_array(ops)
            pw, k4 = _lib.ptr_array(self.W)
            ldw = self.W[0].shape[1]
            if not fused:
                pf, k5 = _lib.ptr_array(self.Fhat)
                ldf = self.Fhat[0].shape[1]
            else:
                pf = k5 = ldf = None

            tr = 1 if self.two_pass_t else 0

            def rows(pi_=pi_, ld=ld, keep=(k3, k4)):      # along y; stored transposed (W[qy][p1]) in two-pass mode
                _lib.check(lib.mlb_fft_rows(pi_, ld, pw, ldw, Rx, Ry, self.sx, self.sy, self.tw2.data_ptr(),
                                            roll_r, roll_c, (h2 // self.sy) % Ry, tr, 4, _stream_ptr()), "mlb_fft_rows")

            def cols(keep=k5):                            # along x
                if tr:                                    # rows of W = fixed qy, contiguous p1 -> Fhat[qx][qy]
                    _lib.check(lib.mlb_fft_rows(pw, ldw, pf, ldf, Ry, Rx, 1, 1, self.tw1.data_ptr(), 0, 0,
                                                (h1 // self.sx) % Rx, 1, 4, _stream_ptr()), "mlb_fft_rows(pass 2)")
                else:
                    _lib.check(lib.mlb_fft_cols(pw, ldw, pf, ldf, Rx, Ry, self.tw1.data_ptr(), (h1 // self.sx) % Rx, 4,
                                                _stream_ptr()), "mlb_fft_cols")
            out.append(("fold_fft_rows" if (self.sx > 1 or self.sy > 1) else "fft_rows", rows,
                        32 * (self.Mx * self.My + Rx * Ry), 4 * 5.0 * Rx * Ry * math.log2(Ry)))
            if fused:
                def cols_power(keep=k4, accumulate=accumulate):   # column pass + radiated power + total_P in one call
                    _lib.check(lib.mlb_fft_cols_power_total(pw, ldw, Rx, Ry, self.tw1.data_ptr(), (h1 // self.sx) % Rx,
                                                            self.d_ux.data_ptr(), self.d_uy.data_ptr(), self.dxp * self.dyp,
                                                            self.wavelength, self.n_glass, Z0, self.P.data_ptr(),
                                                            self.P.shape[1], 1 if accumulate else 0,
                                                            self.block_sums.data_ptr(), self.total.data_ptr(),
                                                            self.dux * self.duy, self.done_counter.data_ptr(),
                                                            _stream_ptr()), "mlb_fft_cols_power_total")
                    return self.P, self.total
                out.append(("fft_cols_power", cols_power, (32 + 4) * Rx * Ry, 4 * 5.0 * Rx * Ry * math.log2(Rx)))
                return out
            out.append(("fft_rows_pass2" if tr else "fft_cols", cols, 64 * Rx * Ry, 4 * 5.0 * Rx * Ry * math.log2(Rx)))
        else:
            pa, k6 = _lib.ptr_array(ops)
            pu, k7 = _lib.ptr_array(self.UT)
            pf, k8 = _lib.ptr_array(self.Fhat)
            ldu, ldf = self.UT[0].shape[1], self.Fhat[0].shape[1]

            def stage1(pa=pa, ld=ld, keep=(k6, k7)):      # UT_f[m2][i] = sum_{m1} J_f[m1][m2] * AxT[m1][i]
                _lib.check(lib.mlb_cgemm_tn(pa, ld, self.AxT.data_ptr(), self.AxT.shape[1], pu, ldu, Ry, Kx, Rx, 4,
                                            _stream_ptr()), "mlb_cgemm_tn(stage 1)")

            def stage2(keep=k8):                          # Fhat_f[i][j] = sum_{m2} UT_f[m2][i] * Ay[m2][j]
                _lib.check(lib.mlb_cgemm_tn(pu, ldu, self.Ay.data_ptr(), self.Ay.shape[1], pf, ldf, Kx, Ky, Ry, 4,
                                            _stream_ptr()), "mlb_cgemm_tn(stage 2)")
            out.append(("cgemm_stage1", stage1, 8 * (4 * Rx * Ry + Rx * Kx + 4 * Ry * Kx), 32.0 * Rx * Ry * Kx))
            out.append(("cgemm_stage2", stage2, 8 * (4 * Ry * Kx + Ry * Ky + 4 * Kx * Ky), 32.0 * Ry * Kx * Ky))
        out.append(("epilogue", self.power, 36 * Kx * Ky, 0.0))
        return out

    def aperture_sums(self, fields):
        """[fold ->] separable reduction or FFT passes:
        Fhat_f[i,j] = sum_{m1,m2} J_f[m1,m2] e^{-ik(x'ux_i + y'uy_j)}.
        Returns the 4 pitched device tensors (logical view [:, :Ky])."""
        for name, fn, _b, _f in self.steps(fields, fused=False)[:-1]:
            fn()
        return self.Fhat

    def power(self, Fhat=None, amp_scale=None, accumulate=False):
        """Epilogue: aperture sums -> P (device), total_P (device scalar, float64).  With
        accumulate=True this run's power is ADDED to P (incoherent sum over sources)."""
        Fhat = self.Fhat if Fhat is None else Fhat
        amp = self.dxp * self.dyp if amp_scale is None else amp_scale
        pf, _k = _lib.ptr_array(Fhat)
        rc = self.lib.mlb_ff_epilogue(pf, Fhat[0].shape[1], self.d_ux.data_ptr(), self.d_uy.data_ptr(),
                                      self.Kx, self.Ky, float(amp), self.wavelength, self.n_glass, Z0,
                                      self.P.data_ptr(), self.P.shape[1],
                                      (1 if self.p_dtype == torch.float64 else 0) + (2 if accumulate else 0),
                                      self.block_sums.data_ptr(), _stream_ptr())
        _lib.check(rc, "mlb_ff_epilogue")
        rc = self.lib.mlb_sum_f64(self.block_sums.data_ptr(), self.nblocks, self.dux * self.duy,
                                  self.total.data_ptr(), _stream_ptr())
        _lib.check(rc, "mlb_sum_f64")
        return self.P, self.total

    def bind_output(self, P):
        """Make run() write the power map straight into the caller's (Kx, Ky) tensor (e.g. a slot of a tile stack that
        is exchanged between GPUs) instead of the plan's own buffer: saves the copy."""
        assert tuple(P.shape) == (self.Kx, self.Ky) and P.dtype == self.p_dtype and P.is_contiguous() and P.is_cuda
        self.P = P

    def run(self, fields, accumulate=False):
        """Device-resident fields (4 CUDA complex64 (Mx,My) tensors) -> (P, total_P) on device.
        accumulate=True adds this item's power to P (total_P is that of this item alone)."""
        if self.fused:
            self._fields_last = fields
            for name, fn, _b, _f in self.steps(fields, accumulate=accumulate):
                fn()
            return self.P, self.total
        self.aperture_sums(fields)
        return self.power(accumulate=accumulate)

    def run_split(self, fields, accumulate=False):
        """run() as two halves for callers that pipeline items over two streams (sharding.py): returns
        (first, second) thunks -- `first` is the HBM-bound pass over the aperture, `second` everything
        after it (column pass, power epilogue, total_P) and returns (P, total_P)."""
        st = self.steps(fields, accumulate=accumulate)
        if self.fused:
            self._fields_last = fields
            return st[0][1], st[1][1]
        head, tail = st[0][1], [fn for _n, fn, _b, _f in st[1:-1]]

        def second():
            for fn in tail:
                fn()
            return self.power(accumulate=accumulate)
        return head, second

    def run_incoherent(self, field_sets):
        """Incoherent sum over several sources / polarisations (the x-, y-, z-dipole recipe of
        nearfield.py:69-73): P = sum_k P_k, total_P = sum_k total_k, everything on the device."""
        total = torch.zeros(1, dtype=torch.float64, device=self.device)
        for k, fields in enumerate(field_sets):
            _, t = self.run(fields, accumulate=(k > 0))
            total += t
        return self.P, total

    def amplitudes(self):
        """Complex aperture sums of the last run as a (4, Kx, Ky) device tensor.  A fused plan keeps them
        in registers during run(); here it re-runs the separate passes on the last run's fields."""
        if self.fused:
            if self._fields_last is None:
                raise _lib.MetalensB200Error("amplitudes(): no run() yet")
            self.aperture_sums(self._fields_last)
        return torch.stack([f[:, :self.Ky] for f in self.Fhat])

    def run_host(self, Ex, Ey=None, Hx=None, Hy=None):
        """End-to-end call with HOST fields: H2D from pinned memory, kernels, D2H.

        Pass four numpy arrays (Mx,My) (converted to complex64 into an internal pinned
        buffer), or one pinned torch complex64 tensor (4,Mx,My) that is copied as is.
        Returns (P numpy, total_P float)."""
        if self._pinned is None:
            self._dev_in = torch.empty((4, self.Mx, _even(self.My)), dtype=torch.complex64, device=self.device)
            # pinned next to this GPU (hostmem.py): the copies of several ranks then do not share an inter-socket link
            self._p_host = hostmem.pinned_empty((self.Kx, self.Ky), self.p_dtype, self.device.index)
            self._t_host = hostmem.pinned_empty((1,), torch.float64, self.device.index)
            self._pinned = True
        if isinstance(Ex, torch.Tensor) and Ey is None:
            src = Ex
            assert src.dtype == torch.complex64 and tuple(src.shape) == (4, self.Mx, self.My)
        else:
            if getattr(self, "_pin_in", None) is None:
                self._pin_in = hostmem.pinned_empty((4, self.Mx, self.My), torch.complex64, self.device.index)
            pin = self._pin_in.numpy()
            for i, a in enumerate((Ex, Ey, Hx, Hy)):
                pin[i] = a                  # dtype conversion to complex64 happens here
            src = self._pin_in
        self._dev_in[:, :, :self.My].copy_(src, non_blocking=True)
        P, total = self.run([self._dev_in[i][:, :self.My] for i in range(4)])
        self._p_host.copy_(P, non_blocking=True)
        self._t_host.copy_(total, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._p_host.numpy().copy(), float(self._t_host[0])

    @property
    def h2d_bytes(self):
        return 4 * self.Mx * self.My * 8

    @property
    def d2h_bytes(self):
        return self.Kx * self.Ky * self.P.element_size() + 8


def farfield_from_fields(Ex, Ey, Hx, Hy, xp_list, yp_list, wavelength, n_glass, stride=1,
                         ux=None, uy=None, method="auto", p_dtype=torch.float64):
    """Real-space aperture fields (host arrays) -> far field, aperture sum included.

    With the default ``stride=1`` the result equals the reference chain
    ``farfield_from_nearfield(fft2(fftshift(Ex)), ...)`` (nearfield_farfield.py:14-75):
    same return tuple ``(P, total_P, ux (Kx,1), uy (1,Ky), dux, duy)``.
    """
    dxp = xp_list[1] - xp_list[0]
    dyp = yp_list[1] - yp_list[0]
    nx, ny = len(xp_list), len(yp_list)
    assert Ex.shape == Ey.shape == Hx.shape == Hy.shape == (nx, ny)
    _check_axis(xp_list, wavelength)
    _check_axis(yp_list, wavelength)
    if ux is not None:
        plan = FarfieldPlan((nx, ny), dxp, dyp, wavelength, n_glass, ux=ux, uy=uy,
                            method=method if method in ("dense", "czt", "tc") else "auto", p_dtype=p_dtype)
    else:
        plan = FarfieldPlan((nx, ny), dxp, dyp, wavelength, n_glass, stride=stride, method=method, p_dtype=p_dtype)
    P, total = plan.run_host(Ex, Ey, Hx, Hy)
    return P, total, plan.ux.reshape(-1, 1), plan.uy.reshape(1, -1), plan.dux, plan.duy


def farfield_from_nearfield(fftEx, fftEy, fftHx, fftHy, xp_list, yp_list, wavelength, n_glass):
    """Drop-in for the reference's ``nearfield_farfield.farfield_from_nearfield``
    (nearfield_farfield.py:14-75): same arguments (``fftEx = fft2(fftshift(Ex))`` ...),
    same return tuple ``(P_here_times_r2_over_uz, total_P, ux, uy, dux, duy)``, same
    AssertionErrors on shape / grid violations.  The power epilogue runs on the GPU in
    float64 on the caller's complex128 values (no down-cast); the reference's RAM chunk loop and
    its progress prints are not reproduced.
    """
    dxp = xp_list[1] - xp_list[0]
    dyp = yp_list[1] - yp_list[0]
    num_x, num_y = len(xp_list), len(yp_list)
    assert fftEx.shape == fftEy.shape == fftHx.shape == fftHy.shape == (num_x, num_y)
    _check_axis(xp_list, wavelength)
    _check_axis(yp_list, wavelength)
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.MetalensB200Error("metalens_b200 needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    ux = fft_bin_direction_cosines(num_x, dxp, wavelength, n_glass)
    uy = fft_bin_direction_cosines(num_y, dyp, wavelength, n_glass)
    # the caller's complex128 arrays as they are: the whole epilogue (:135-189) runs in float64 on float64 inputs
    F = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)).to(dev) for a in (fftEx, fftEy, fftHx, fftHy)]
    d_ux, d_uy = torch.from_numpy(ux).to(dev), torch.from_numpy(uy).to(dev)
    P = torch.empty((num_x, num_y), dtype=torch.float64, device=dev)
    ux_s = np.fft.fftshift(ux)                                                           # :69
    uy_s = np.fft.fftshift(uy)                                                           # :70
    dux = ux_s[1] - ux_s[0]                                                              # :71
    duy = uy_s[1] - uy_s[0]                                                              # :72
    nblocks = lib.mlb_ff_epilogue_blocks(num_x, num_y)
    block_sums = torch.empty(nblocks, dtype=torch.float64, device=dev)
    total = torch.zeros(1, dtype=torch.float64, device=dev)
    pf, _keep = _lib.ptr_array(F)
    _lib.check(lib.mlb_ff_epilogue(pf, num_y, d_ux.data_ptr(), d_uy.data_ptr(), num_x, num_y, float(dxp * dyp),
                                   float(wavelength), float(n_glass), Z0, P.data_ptr(), num_y, 1 + 4,
                                   block_sums.data_ptr(), _stream_ptr()), "mlb_ff_epilogue")
    _lib.check(lib.mlb_sum_f64(block_sums.data_ptr(), nblocks, float(dux * duy), total.data_ptr(), _stream_ptr()),
               "mlb_sum_f64")                                                            # :74
    P = torch.roll(P, shifts=(num_x // 2, num_y // 2), dims=(0, 1)).cpu().numpy()       # fftshift, :68
    ux2, uy2 = np.meshgrid(ux_s, uy_s, indexing="ij", sparse=True)                       # :73
    return P, float(total.cpu()[0]), ux2, uy2, dux, duy
