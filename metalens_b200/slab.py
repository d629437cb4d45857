"""ONE aperture, far field spread over several GPUs (BASELINE config 4).

The reference computes disjoint uy chunks of one transform separately (nearfield_farfield.py:45-66) and
disjoint y slabs of one assembly separately (nearfield.py:488-514).  Here the same independence is cut the
other way round so that BOTH halves of the hot path shrink by 1/G and one small exchange remains:

    rank g owns every G-th block of 4 folded rows r (round-robin: equal work on a round lens) -- i.e. the aperture
    x-rows { r + t K1 } -- and
      1. assembles only those rows of the aperture   (NearfieldPlan.run on x_pts[x_rows]),
      2. folds + row-transforms them                 (mlb_fft_rows_scatter), the kernel storing each row's column
         slab p straight into rank p's buffer over NVLink: the all-to-all of a distributed 2-D FFT, fused,
      3. one flag barrier                            (mlb_peer_barrier),
      4. column pass + radiated power on ITS K2/G columns (mlb_fft_cols_power) into its slab of P,
      5. ONE all-gather of the P slabs at the end    (mlb_peer_allgather: pushed, with the total_P block sums).

Every kernel is the single-GPU one on a subset of rows / columns, so P is bit-identical to the single-GPU result
(total_P too: the gathered block sums are the single-GPU block sums).  Bytes over the links per rank:
32 K1 K2 (G-1)/G^2 for the all-to-all and 4 K1 K2 (G-1)/G for the gather (15 + 15 MB at K = 2048, G = 8).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .farfield import fft_bin_direction_cosines, _even
from .peer import PeerChannel
from .units import Z0


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


ROW_BLOCK = 4      # rows per ownership block: the x extent of the assembly kernel's warp tile


def _row_block(K1, world):
    per = K1 // world
    return ROW_BLOCK if per % ROW_BLOCK == 0 else 1


def slab_rows(Mx, sx, rank, world):
    """Aperture x indices rank `rank` owns, in the order its local buffer holds them: the sx aliased copies of
    its K1/world folded rows one after the other.  The folded rows are dealt to the ranks round-robin in blocks of
    ROW_BLOCK (rank g owns rows 4(g + G b) .. 4(g + G b) + 3, b = 0, 1, ...): every rank then sees the same mix of
    centre, ring and outside-the-lens samples of a round lens (contiguous slabs cost the edge ranks half as much)."""
    K1 = Mx // sx
    per = K1 // world
    blk = _row_block(K1, world)
    b = np.arange(per // blk)
    base = (blk * (rank + world * b))[:, None] + np.arange(blk)[None, :]
    base = base.reshape(-1)
    return np.concatenate([base + t * K1 for t in range(sx)])


def slab_geometry(Mx, My, sx, sy, rank, world):
    """Index bookkeeping of one rank (pure integers; emulated in numpy by tests/test_slab_index.py).
    With K1 = Mx/sx, K2 = My/sy, h = M//2 (the fftshift origin of nearfield_farfield.py:18-20):
      folded row  G[r][p]  = sum_{t1,t2} J[r + t1 K1][((p - roll_c) mod K2) + t2 K2]      r = this rank's source rows
      row pass    W[R][(q + out_roll_rows) mod K2] = sum_p G[r][p] e^{-2 pi i q p / K2},   R = (r + h1) mod K1
      column pass F[(q + out_roll_cols) mod K1][c] = sum_R W[R][c] e^{-2 pi i q R / K1}
    so that F is in the fftshifted order of ux[::sx], uy[::sy].  The rank's l-th local row is source row
    row_block (rank + world (l // row_block)) + l % row_block, i.e. intermediate row
    (out_row0 + (l // row_block) row_stride + l % row_block) mod K1."""
    K1, K2 = Mx // sx, My // sy
    h1, h2 = Mx // 2, My // 2
    per = K1 // world
    blk = _row_block(K1, world)
    return dict(K1=K1, K2=K2, rows_per_rank=per, cols_per_rank=K2 // world, x_rows=slab_rows(Mx, sx, rank, world),
                roll_c=h2 % K2, out_roll_rows=(h2 // sy) % K2, out_roll_cols=(h1 // sx) % K1,
                out_row0=(blk * rank + h1 % K1) % K1, row_block=blk, row_stride=blk * world)


class SlabFarfield:
    """Far field of one (Mx, My) aperture on the every-`stride`-th-FFT-bin grid, computed by `world` ranks.

    peers : a SymmetricPeers (one process per GPU) or a VirtualPeers.view(rank) (virtual ranks on one device).
    gather_ctas: CTAs of the pushed all-gather of P (default 128 across real GPUs: it is on the critical path with nothing
    else running, so it gets most of the SMs; 8 for virtual ranks, whose waiting kernels share one GPU).
    run(fields) takes THIS rank's rows of the four fields -- complex64 (len(x_rows), My) each, row order
    ``self.x_rows`` -- and returns (P (K1, K2) float32 complete on every rank, total_P device scalar).
    """

    def __init__(self, shape, dxp, dyp, wavelength, n_glass, stride, peers, name="slab", gather_ctas=None):
        self.lib = lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MetalensB200Error("metalens_b200 needs a CUDA device (no CPU fallback)")
        self.peers = peers
        self.rank, self.world = peers.rank, peers.world
        dev = self.device = peers.device
        self.Mx, self.My = int(shape[0]), int(shape[1])
        self.dxp, self.dyp = float(dxp), float(dyp)
        self.wavelength, self.n_glass = float(wavelength), float(n_glass)
        sx, sy = (stride, stride) if np.isscalar(stride) else stride
        self.sx, self.sy = int(sx), int(sy)
        G = self.world
        assert self.Mx % self.sx == 0 and self.My % self.sy == 0
        K1, K2 = self.Mx // self.sx, self.My // self.sy
        self.K1, self.K2 = K1, K2
        ok = ((self.sx == 1 or (self.Mx // 2) % self.sx == 0) and (self.sy == 1 or (self.My // 2) % self.sy == 0)
              and K1 % G == 0 and K2 % G == 0 and (G & (G - 1)) == 0 and (K2 // G) % 4 == 0
              and (K2 & (K2 - 1)) == 0 and 256 <= K2 <= 2048 and lib.mlb_fft_cols_power_blocks(K1, K2 // G) > 0)
        if not ok:
            raise ValueError("SlabFarfield: needs a power-of-two rank count dividing the folded sizes, folded row length "
                             "a power of two in 256..2048 and a folded column length a power of two in 256..8192 "
                             "(got %d x %d over %d ranks)" % (K1, K2, G))
        self.rows_per_rank, self.cols_per_rank = K1 // G, K2 // G
        self.x_rows = slab_rows(self.Mx, self.sx, self.rank, G)
        self.ux = np.ascontiguousarray(np.fft.fftshift(fft_bin_direction_cosines(self.Mx, self.dxp, wavelength, n_glass))[::self.sx])
        self.uy = np.ascontiguousarray(np.fft.fftshift(fft_bin_direction_cosines(self.My, self.dyp, wavelength, n_glass))[::self.sy])
        self.dux, self.duy = float(self.ux[1] - self.ux[0]), float(self.uy[1] - self.uy[0])
        self.d_ux = torch.from_numpy(self.ux).to(dev)
        self.d_uy = torch.from_numpy(self.uy).to(dev)
        self.tw1 = torch.empty(2 * K1, dtype=torch.complex64, device=dev)
        self.tw2 = torch.empty(2 * K2, dtype=torch.complex64, device=dev)
        for t, n in ((self.tw1, K1), (self.tw2, K2)):
            _lib.check(lib.mlb_fft_twiddle(n, t.data_ptr(), _stream_ptr()), "mlb_fft_twiddle")
        # peer-visible buffers: the intermediate (all K1 rows of this rank's K2/G columns, 4 fields), the full P,
        # the block sums of total_P
        self.ldw = _even(self.cols_per_rank)
        w_bytes = 4 * K1 * self.ldw * 8
        self._W, w_ptrs = peers.alloc(name + ".W", w_bytes)
        self.W = self._W[:w_bytes].view(torch.complex64).view(4, K1, self.ldw)
        self._w_ptrs = [[p + f * K1 * self.ldw * 8 for f in range(4)] for p in w_ptrs]       # [peer][field]
        self._P, self._p_ptrs = peers.alloc(name + ".P", K1 * K2 * 4)
        self.P = self._P[:K1 * K2 * 4].view(torch.float32).view(K1, K2)
        self.nb_local = lib.mlb_fft_cols_power_blocks(K1, self.cols_per_rank)
        self.nb_total = lib.mlb_fft_cols_power_blocks(K1, K2)
        # the gathered block sums equal the single-GPU ones when a CTA's blocks are contiguous per column tile
        self.exact_total = (self.nb_local * G == self.nb_total and K1 <= 2048)
        self._bs, self._bs_ptrs = peers.alloc(name + ".bs", 8 * max(self.nb_local * G, 2))
        self.block_sums = self._bs[:8 * self.nb_local * G].view(torch.float64)
        self.total = torch.zeros(1, dtype=torch.float64, device=dev)
        self.chan = PeerChannel(peers, name + ".chan")
        if gather_ctas is None:          # virtual ranks share one GPU: their spinning CTAs must leave room for each other
            gather_ctas = 8 if getattr(peers, "virtual", False) else 128
        self.gather_ctas = gather_ctas
        peers.sync()
        # fftshift bookkeeping of nearfield_farfield.py:18-20, :68 as index rolls
        geo = slab_geometry(self.Mx, self.My, self.sx, self.sy, self.rank, G)
        self.roll_c, self.out_roll_rows = geo["roll_c"], geo["out_roll_rows"]
        self.out_roll_cols, self.out_row0 = geo["out_roll_cols"], geo["out_row0"]
        self.row_block, self.row_stride = geo["row_block"], geo["row_stride"]
        self._pw_all, self._k1 = _lib.ptr_array([p for peer in self._w_ptrs for p in peer])
        self._pw_mine, self._k2 = _lib.ptr_array(self._w_ptrs[self.rank])

    # ------------------------------------------------------------------
    def _rows(self, fields):
        lib = self.lib
        n_loc = self.x_rows.size
        for f in fields:
            assert f.is_cuda and f.dtype == torch.complex64 and tuple(f.shape) == (n_loc, self.My)
            assert f.stride(1) == 1 and f.stride(0) % 2 == 0 and f.data_ptr() % 16 == 0
        ld = fields[0].stride(0)
        assert all(f.stride(0) == ld for f in fields)
        pin, keep = _lib.ptr_array(list(fields))
        _lib.check(lib.mlb_fft_rows_scatter(pin, ld, self._pw_all, self.ldw, self.rows_per_rank, self.K2, self.sx, self.sy,
                                            self.tw2.data_ptr(), self.roll_c, self.out_roll_rows, self.out_row0,
                                            self.row_block, self.row_stride, self.K1, self.world, 4, _stream_ptr()),
                   "mlb_fft_rows_scatter")

    def _cols(self):
        col0 = self.rank * self.cols_per_rank
        bs0 = self.rank * self.nb_local
        _lib.check(self.lib.mlb_fft_cols_power(self._pw_mine, self.ldw, self.K1, self.cols_per_rank, self.tw1.data_ptr(),
                                               self.out_roll_cols, self.d_ux.data_ptr(), self.d_uy.data_ptr() + 8 * col0,
                                               self.dxp * self.dyp, self.wavelength, self.n_glass, Z0,
                                               self.P.data_ptr() + 4 * col0, self.K2, 0,
                                               self.block_sums.data_ptr() + 8 * bs0, None, 0, _stream_ptr()),
                   "mlb_fft_cols_power")

    def warm(self, fields):
        """Launch the compute kernels once without any exchange (loads them: CUDA loads kernels lazily, and a first
        launch may have to wait for running kernels -- fatal for virtual ranks whose kernels wait for each other)."""
        self._rows(fields)
        self._cols()
        self._sum()
        torch.cuda.current_stream().synchronize()

    def run(self, fields, wait=True):
        """fields: 4 CUDA complex64 tensors (len(x_rows), My) sharing one even row pitch.  Launches the five steps
        on the current stream.  wait=False leaves out the final mlb_peer_wait (call finish() before reading P)."""
        self._rows(fields)
        self.chan.barrier()
        self._cols()
        if self.world > 1:
            col0 = self.rank * self.cols_per_rank
            bs0 = self.rank * self.nb_local
            self.chan.allgather(self.P.data_ptr() + 4 * col0, 4 * self.K2, self.K1, 4 * self.cols_per_rank,
                                self._p_ptrs, 4 * self.K2, 4 * col0,
                                aux=(self.block_sums.data_ptr() + 8 * bs0, self._bs_ptrs, bs0, self.nb_local),
                                n_ctas=self.gather_ctas)
            if wait:
                self.finish()
                return self.P, self.total
            return self.P, None
        self._sum()
        return self.P, self.total

    def _sum(self):
        _lib.check(self.lib.mlb_sum_f64(self.block_sums.data_ptr(), self.nb_local * self.world, self.dux * self.duy,
                                        self.total.data_ptr(), _stream_ptr()), "mlb_sum_f64")

    def finish(self):
        """Every rank's slab of P (and block sums) has arrived; total_P summed in the single-GPU order."""
        if self.world > 1:
            self.chan.wait_sum(self.block_sums, self.nb_local * self.world, self.dux * self.duy, self.total)
        else:
            self._sum()
        return self.P, self.total


def assemble_slab(nf_plan, slab, source, source_pol, x_pts, y_pts, dipole_moment=1e-30, out=None, check=True):
    """Hot path B for a SlabFarfield rank: the fused assembly kernel on this rank's aperture rows only
    (nearfield.py:488-514 builds disjoint slabs independently).  Returns (fields (4, rows, ld) complex64,
    partial incident power (device scalar, sum over ranks = the full lens))."""
    x_pts = np.asarray(x_pts, dtype=np.float64)
    y_pts = np.asarray(y_pts, dtype=np.float64)
    dxdy = float(x_pts[1] - x_pts[0]) * float(y_pts[1] - y_pts[0])
    return nf_plan.run(source[0], source[1], source[2], source_pol, x_pts[slab.x_rows], y_pts,
                       dipole_moment=dipole_moment, out=out, dxdy=dxdy, check=check)
