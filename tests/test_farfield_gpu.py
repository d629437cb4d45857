"""GPU parity tests for hot path A (NF->FF) through the C-ABI.

Oracle = outputs of the unmodified reference (tests/golden) and oracle/farfield_oracle.py.
Tolerance: north_star's 1e-5 (max|dP|/max|P_ref| over finite bins, identical NaN masks).
"""
import os

import numpy as np
import pytest

import apertures
from parity import FF_TOL, field_error, power_map_error

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

WL = 532e-9
NG = apertures.N_GLASS[532]

CASES = {
    "kat1_uniform": lambda: apertures.uniform(128, WL, NG),
    "kat2_disc": lambda: apertures.disc(128, WL, NG),
    "kat3_tilted": lambda: apertures.tilted_te(128, WL, NG),
    "rand128_seed0": lambda: apertures.gaussian_random(128, 0, WL),
    "rand_48x40_seed5": lambda: apertures.gaussian_random(48, 5, WL, My=40),
    "rand_45x27_seed6": lambda: apertures.gaussian_random(45, 6, WL, My=27),
    "lens256_seed1": lambda: apertures.focusing_lens(256, 1, WL, NG),
    "lens256_seed1_rot": lambda: apertures.focusing_lens(256, 1, WL, NG, rotate=True),
}


def golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, "farfield_%s.npz" % name))


@pytest.mark.parametrize("name", sorted(CASES))
def test_dropin_farfield_from_nearfield(name, golden_dir):
    """Reference signature: FFT'd fields in, reference return tuple out."""
    from metalens_b200.farfield import farfield_from_nearfield
    g = golden(golden_dir, name)
    Ex, Ey, Hx, Hy, x, y = CASES[name]()
    f = [np.fft.fft2(np.fft.fftshift(a.astype(complex))) for a in (Ex, Ey, Hx, Hy)]
    P, total, ux, uy, dux, duy = farfield_from_nearfield(f[0], f[1], f[2], f[3], list(x), list(y), WL, NG)
    assert P.dtype == np.float64 and P.shape == g["P"].shape
    # float64 epilogue on the caller's complex128 arrays: far below north_star's 1e-5
    assert power_map_error(P, g["P"]) < 1e-11
    assert abs(total - g["total_P"]) <= 1e-11 * abs(g["total_P"])
    assert ux.shape == (len(x), 1) and uy.shape == (1, len(y))
    np.testing.assert_array_equal(ux.ravel(), g["ux"])
    np.testing.assert_array_equal(uy.ravel(), g["uy"])
    assert dux == g["dux"] and duy == g["duy"]


@pytest.mark.parametrize("name,rot", [("lens675_seed9", False), ("lens675_seed9_rot", True)])
def test_reference_default_size_675_against_reference(name, rot, golden_dir):
    """675 x 675 (the reference's own typical grid: good_fft_number, odd, 3^3 5^2 -> mixed-radix FFT passes, odd-row
    epilogue) against a sampled subset of the UNMODIFIED reference's map; and the strict drop-in on the same case."""
    from metalens_b200.farfield import farfield_from_fields, farfield_from_nearfield
    g = golden(golden_dir, name)
    wl, ng = float(g["wavelength"]), float(g["n_glass"])
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(675, 9, wl, ng, rotate=rot)
    P, total, ux, uy, dux, duy = farfield_from_fields(Ex, Ey, Hx, Hy, x, y, wl, ng, stride=1, method="fft",
                                                      p_dtype=torch.float32)
    got = P.ravel()[g["index"]]
    assert np.array_equal(np.isnan(got), np.isnan(g["P"])) and int(np.isnan(P).sum()) == int(g["nan_count"])
    fin = np.isfinite(g["P"])
    assert np.abs(got - g["P"])[fin].max() / float(g["P_max"]) < FF_TOL
    assert abs(total - float(g["total_P"])) <= FF_TOL * abs(float(g["total_P"]))
    np.testing.assert_array_equal(ux.ravel(), g["ux"])
    f = [np.fft.fft2(np.fft.fftshift(a.astype(complex))) for a in (Ex, Ey, Hx, Hy)]
    P2, total2, *_ = farfield_from_nearfield(f[0], f[1], f[2], f[3], list(x), list(y), wl, ng)
    got2 = P2.ravel()[g["index"]]
    assert np.array_equal(np.isnan(got2), np.isnan(g["P"]))
    assert np.abs(got2 - g["P"])[fin].max() / float(g["P_max"]) < 1e-11
    assert abs(total2 - float(g["total_P"])) <= 1e-11 * abs(float(g["total_P"]))


@pytest.mark.parametrize("method", ["auto", "dense"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_fields_to_farfield_all_bins(name, method, golden_dir):
    """Aperture sum on the GPU on the reference's full FFT-bin grid: shared-memory FFT passes for
    power-of-two apertures ('auto'), the dense tiled reduction otherwise and when forced,
    including odd and non-square apertures."""
    from metalens_b200.farfield import farfield_from_fields
    g = golden(golden_dir, name)
    Ex, Ey, Hx, Hy, x, y = CASES[name]()
    P, total, ux, uy, dux, duy = farfield_from_fields(Ex, Ey, Hx, Hy, x, y, WL, NG, stride=1, method=method)
    assert power_map_error(P, g["P"]) < FF_TOL
    assert abs(total - g["total_P"]) <= FF_TOL * abs(g["total_P"])
    np.testing.assert_array_equal(ux.ravel(), g["ux"])


@pytest.mark.parametrize("method", ["dense", "fold", "fft"])
@pytest.mark.parametrize("name,stride", [("rand128_seed0", 4), ("lens256_seed1", 8), ("lens256_seed1_rot", 2),
                                         ("rand_48x40_seed5", (4, 2))])
def test_strided_bins(name, stride, method, golden_dir):
    """BASELINE configs 2/3 shape: K < M, far-field grid = every s-th fftshifted reference bin."""
    from metalens_b200.farfield import farfield_from_fields
    g = golden(golden_dir, name)
    Ex, Ey, Hx, Hy, x, y = CASES[name]()
    sx, sy = (stride, stride) if np.isscalar(stride) else stride
    # (folded size 12 x 20 of the 48 x 40 case is not a power of two: mixed-radix passes)
    P, total, ux, uy, dux, duy = farfield_from_fields(Ex, Ey, Hx, Hy, x, y, WL, NG, stride=stride, method=method,
                                                      p_dtype=torch.float32)
    ref = g["P"][::sx, ::sy]
    assert P.dtype == np.float32
    assert power_map_error(P, ref) < FF_TOL
    ref_total = ref[np.isfinite(ref)].sum() * dux * duy
    assert abs(total - ref_total) <= FF_TOL * abs(ref_total)


def test_arbitrary_direction_cosine_grid():
    """A4: direct sum onto a grid that is NOT a subset of the FFT bins."""
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import farfield_from_fields
    Ex, Ey, Hx, Hy, x, y = apertures.gaussian_random(96, 11, WL, My=70)
    ux = np.linspace(-0.83, 0.91, 53)
    uy = np.linspace(-0.5, 0.77, 38)
    ux[7] = 0.0
    uy[5] = 0.0                                  # exercise the exact-DC special case
    P, total, *_ = farfield_from_fields(Ex, Ey, Hx, Hy, x, y, WL, NG, ux=ux, uy=uy)
    P_ref, _ = fo.farfield_dense(Ex, Ey, Hx, Hy, x[1] - x[0], y[1] - y[0], ux, uy, WL, NG)
    assert power_map_error(P, P_ref) < FF_TOL


@pytest.mark.parametrize("shape,kx,ky,span", [((96, 70), 53, 38, (-0.83, 0.91, -0.5, 0.77)),       # wide, non-square
                                              ((128, 128), 64, 64, (-0.06, 0.06, -0.05, 0.07)),    # zoom around the axis
                                              ((45, 27), 30, 9, (0.2, -0.3, 0.1, 0.4)),            # odd sizes, descending ux
                                              ((1024, 1024), 300, 301, (-0.02, 0.02, -0.02, 0.02))])
def test_chirp_z_uniform_zoom_grids(shape, kx, ky, span):
    """A4 on a UNIFORM grid that is not a set of FFT bins: method 'czt' (chirp-z on the FFT passes) against the float64
    direct sum, and against the tiled reduction it replaces; 'auto' picks it for explicit uniform grids."""
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import FarfieldPlan
    Mx, My = shape
    if Mx >= 1024:
        Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(Mx, 21, WL, NG, na=0.3)
    else:
        Ex, Ey, Hx, Hy, x, y = apertures.gaussian_random(Mx, 21, WL, My=My)
    ux = np.linspace(span[0], span[1], kx)
    uy = np.linspace(span[2], span[3], ky)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (Ex, Ey, Hx, Hy)]
    dx, dy = x[1] - x[0], y[1] - y[0]
    plan = FarfieldPlan((Mx, My), dx, dy, WL, NG, ux=ux, uy=uy)
    assert plan.method == "czt"
    P, total = plan.run(dev)
    P = P.cpu().numpy()
    P_ref, F_ref = fo.farfield_dense(Ex, Ey, Hx, Hy, dx, dy, ux, uy, WL, NG)
    assert power_map_error(P, P_ref) < FF_TOL
    amps = plan.amplitudes().cpu().numpy()
    for k in range(4):
        assert field_error(amps[k], F_ref[k]) < 3e-6
    dense = FarfieldPlan((Mx, My), dx, dy, WL, NG, ux=ux, uy=uy, method="dense")
    Pd, td = dense.run(dev)
    assert power_map_error(Pd.cpu().numpy(), P_ref) < FF_TOL
    assert abs(total.item() - td.item()) <= FF_TOL * abs(td.item())
    # a slab of far-field rows (multi-GPU tile) is still a uniform grid
    slab = FarfieldPlan((Mx, My), dx, dy, WL, NG, ux=ux, uy=uy, rows=(kx // 3, kx // 3 + 7))
    assert slab.method == "czt"
    Ps = slab.run(dev)[0].cpu().numpy()
    assert power_map_error(Ps, P_ref[kx // 3:kx // 3 + 7]) < FF_TOL * max(1.0, np.nanmax(P_ref) / np.nanmax(P_ref[kx // 3:kx // 3 + 7]))


def test_chirp_z_on_the_strided_bin_grid():
    """The every-4th-bin grid of BASELINE cfg 2/3 is uniform too: 'czt' must agree with the fold + FFT path there."""
    from metalens_b200.farfield import FarfieldPlan
    M, s = 512, 4
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 5, WL, NG)
    dev = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    d = x[1] - x[0]
    a = FarfieldPlan((M, M), d, d, WL, NG, stride=s, method="fft")
    b = FarfieldPlan((M, M), d, d, WL, NG, stride=s, method="czt")
    Pa, ta = a.run(dev)
    Pb, tb = b.run(dev)
    assert power_map_error(Pb.cpu().numpy(), Pa.cpu().numpy()) < FF_TOL
    assert abs(ta.item() - tb.item()) <= FF_TOL * abs(ta.item())
    with pytest.raises(ValueError):
        FarfieldPlan((8192, 8192), d, d, WL, NG, stride=4, method="czt")          # M + K - 1 > 8192


def test_complex_amplitudes_match_fft():
    """The aperture sums themselves (not only P) equal fft2(fftshift(.)) -- phase origin (Q4)."""
    from metalens_b200.farfield import FarfieldPlan
    for M, My, seed, method in ((64, 64, 3, "dense"), (45, 27, 6, "dense"), (64, 64, 3, "fft"), (32, 256, 8, "fft")):
        fields = apertures.gaussian_random(M, seed, WL, My=My)
        x, y = fields[4], fields[5]
        plan = FarfieldPlan((M, My), x[1] - x[0], y[1] - y[0], WL, NG, stride=1, method=method)
        dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in fields[:4]]
        plan.run(dev)
        amps = plan.amplitudes().cpu().numpy()
        for k in range(4):
            ref = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(fields[k].astype(complex))))
            assert field_error(amps[k], ref) < 2e-6


def test_row_slabs_equal_full_far_field(golden_dir):
    """Multi-GPU tile = a slab of far-field rows: slabs computed separately (as different ranks
    would) reassemble to the single-plan result bit for bit (dense and fold)."""
    from metalens_b200.farfield import FarfieldPlan
    Ex, Ey, Hx, Hy, x, y = CASES["lens256_seed1"]()
    dev = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    d = x[1] - x[0]
    for method in ("dense", "fold"):
        full = FarfieldPlan((256, 256), d, d, WL, NG, stride=4, method=method)
        P_full = full.run(dev)[0].clone()
        parts = []
        for r0, r1 in ((0, 16), (16, 32), (32, 64)):
            slab = FarfieldPlan((256, 256), d, d, WL, NG, stride=4, method=method, rows=(r0, r1))
            assert slab.Kx == r1 - r0
            parts.append(slab.run(dev)[0].clone())
        assert torch.equal(torch.cat(parts, dim=0), P_full) or \
            bool(((torch.cat(parts, dim=0) == P_full) | (torch.isnan(P_full))).all())
    with pytest.raises(ValueError):
        FarfieldPlan((256, 256), d, d, WL, NG, stride=4, method="fft", rows=(0, 16))
    with pytest.raises(ValueError):
        FarfieldPlan((154, 256), d, d, WL, NG, stride=1, method="fft")          # 154 = 2*7*11 is not 5-smooth


def test_cgemm_tn_ragged_shapes():
    """C[r][c] = sum_k At[k][r] B[k][c] against torch for ragged sizes, both tile configs."""
    import ctypes as C
    from metalens_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(0)
    for rows, cols, depth, batch in ((1, 1, 1, 1), (33, 17, 5, 2), (130, 257, 77, 4), (640, 1280, 100, 4),
                                     (129, 64, 16, 3)):
        lda, ldb, ldc = rows + (rows & 1) + 2, cols + (cols & 1), cols + (cols & 1) + 4
        At = [torch.randn(depth, lda, dtype=torch.complex64, generator=g).cuda() for _ in range(batch)]
        B = torch.randn(depth, ldb, dtype=torch.complex64, generator=g).cuda()
        Cs = [torch.full((rows, ldc), 7 + 7j, dtype=torch.complex64).cuda() for _ in range(batch)]
        pa, k1 = _lib.ptr_array(At)
        pc, k2 = _lib.ptr_array(Cs)
        rc = lib.mlb_cgemm_tn(pa, lda, B.data_ptr(), ldb, pc, ldc, rows, cols, depth, batch,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "mlb_cgemm_tn")
        torch.cuda.synchronize()
        for b in range(batch):
            ref = (At[b][:, :rows].to(torch.complex128).T @ B[:, :cols].to(torch.complex128))
            got = Cs[b][:, :cols].to(torch.complex128)
            err = (got - ref).abs().max().item() / ref.abs().max().item()
            assert err < 5e-6, (rows, cols, depth, b, err)
            assert torch.all(Cs[b][:, cols:] == 7 + 7j)          # pitch padding untouched


def test_cgemm_rejects_bad_arguments():
    import ctypes as C
    from metalens_b200 import _lib
    lib = _lib.load()
    A = [torch.zeros(4, 3, dtype=torch.complex64).cuda()]
    pa, k1 = _lib.ptr_array(A)
    rc = lib.mlb_cgemm_tn(pa, 3, A[0].data_ptr(), 3, pa, 3, 3, 3, 4, 1, None)
    assert rc != 0 and b"even" in lib.mlb_last_error()
    with pytest.raises(_lib.MetalensB200Error):
        _lib.check(rc, "mlb_cgemm_tn")


def test_fold_matches_numpy():
    import ctypes as C
    from metalens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(4)
    for (M1, M2, s1, s2) in ((16, 24, 4, 2), (30, 21, 3, 7), (64, 64, 1, 4)):
        K1, K2 = M1 // s1, M2 // s2
        h1, h2 = M1 // 2, M2 // 2
        J = [(rng.standard_normal((M1, M2)) + 1j * rng.standard_normal((M1, M2))).astype(np.complex64) for _ in range(4)]
        ldj, ldg = M2 + (M2 & 1), K2 + (K2 & 1)
        dJ = [torch.zeros(M1, ldj, dtype=torch.complex64).cuda() for _ in range(4)]
        for d, a in zip(dJ, J):
            d[:, :M2].copy_(torch.from_numpy(a))
        dG = [torch.zeros(K1, ldg, dtype=torch.complex64).cuda() for _ in range(4)]
        pj, k1 = _lib.ptr_array(dJ)
        pg, k2 = _lib.ptr_array(dG)
        rc = lib.mlb_fold(pj, ldj, M1, M2, s1, s2, h1, h2, pg, ldg, 4, None)
        _lib.check(rc, "mlb_fold")
        torch.cuda.synchronize()
        for a, d in zip(J, dG):
            r = np.roll(a.astype(complex), (h1, h2), axis=(0, 1))      # r[p] = a[p - h]
            ref = r.reshape(s1, K1, s2, K2).sum(axis=(0, 2))
            assert field_error(d[:, :K2].cpu().numpy(), ref) < 1e-6


@pytest.fixture
def fft_engines(request):
    """(rows_engine, cols_engine) for one test; the library defaults (2, 1) are restored afterwards."""
    from metalens_b200 import _lib
    lib = _lib.load()
    rows, cols = request.param
    lib.mlb_set_option(b"rows_engine", rows)
    lib.mlb_set_option(b"cols_engine", cols)
    lib.mlb_set_option(b"r16_min_lg", 8)            # radix-16 kernels from 256 points (default: from 1024)
    yield request.param
    lib.mlb_set_option(b"rows_engine", 2)
    lib.mlb_set_option(b"cols_engine", 1)
    lib.mlb_set_option(b"r16_min_lg", 10)


@pytest.mark.parametrize("fft_engines", [(0, 0), (2, 1)], indirect=True, ids=["radix4", "radix16"])
def test_fft_passes_match_numpy(fft_engines):
    """mlb_fft_rows / mlb_fft_cols against numpy.fft for every power of two up to the limit,
    including the fftshift rolls; once with the radix-4 shared-memory kernels everywhere and once with the
    default engines (radix-16 register kernels from 256 points up)."""
    from metalens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(9)
    assert lib.mlb_fft_max_length() == 8192
    for N, other in ((2, 5), (4, 3), (8, 9), (16, 4), (64, 33), (128, 7), (512, 6), (1024, 5), (2048, 3), (4096, 19),
                     (8192, 2)):
        a = [(rng.standard_normal((other, N)) + 1j * rng.standard_normal((other, N))).astype(np.complex64) for _ in range(2)]
        tw = torch.empty(2 * N, dtype=torch.complex64).cuda()
        _lib.check(lib.mlb_fft_twiddle(N, tw.data_ptr(), None), "tw")
        ld = N + 2
        din = [torch.zeros(other, ld, dtype=torch.complex64).cuda() for _ in a]
        for d, v in zip(din, a):
            d[:, :N].copy_(torch.from_numpy(v))
        dout = [torch.zeros(other, ld, dtype=torch.complex64).cuda() for _ in a]
        rr, rc, ro = other // 2, N // 2, (N // 2 + 1) % N
        pi_, k1 = _lib.ptr_array(din)
        po, k2 = _lib.ptr_array(dout)
        _lib.check(lib.mlb_fft_rows(pi_, ld, po, ld, other, N, 1, 1, tw.data_ptr(), rr, rc, ro, 0, 2, None), "rows")
        torch.cuda.synchronize()
        for v, d in zip(a, dout):
            ref = np.roll(np.fft.fft(np.roll(v.astype(complex), (rr, rc), axis=(0, 1)), axis=1), ro, axis=1)
            assert field_error(d[:, :N].cpu().numpy(), ref) < 3e-6, ("rows", N)
        # columns: transform along axis 0 of the transposed data, in place
        ldc = other + (other & 1) + 2
        dcol = [torch.zeros(N, ldc, dtype=torch.complex64).cuda() for _ in a]
        for d, v in zip(dcol, a):
            d[:, :other].copy_(torch.from_numpy(v.T.copy()))
        pc, k3 = _lib.ptr_array(dcol)
        if N >= 4096:       # four-step decomposition: input is scratch, output must be a different buffer
            assert lib.mlb_fft_cols(pc, ldc, pc, ldc, N, other, tw.data_ptr(), ro, 2, None) != 0
            dco = [torch.zeros(N, ldc, dtype=torch.complex64).cuda() for _ in a]
            pco, k4 = _lib.ptr_array(dco)
            _lib.check(lib.mlb_fft_cols(pc, ldc, pco, ldc, N, other, tw.data_ptr(), ro, 2, None), "cols")
            dcol = dco
        else:
            _lib.check(lib.mlb_fft_cols(pc, ldc, pc, ldc, N, other, tw.data_ptr(), ro, 2, None), "cols")
        torch.cuda.synchronize()
        for v, d in zip(a, dcol):
            ref = np.roll(np.fft.fft(v.astype(complex).T, axis=0), ro, axis=0)
            assert field_error(d[:, :other].cpu().numpy(), ref) < 3e-6, ("cols", N)
    bad = torch.zeros(4, 14, dtype=torch.complex64).cuda()
    pb, k4 = _lib.ptr_array([bad])
    assert lib.mlb_fft_rows(pb, 14, pb, 14, 4, 14, 1, 1, bad.data_ptr(), 0, 0, 0, 0, 1, None) != 0   # 14 = 2*7: not 5-smooth
    # mixed radix (good_fft_number sizes): rows with fold + rolls, and columns
    # engine 1 = big-radix kernels (fftmix.cuh; long columns in two passes when >= 8 columns)
    cases = [(1, c) for c in ((6, 5, 1, 1), (12, 7, 2, 3), (45, 9, 1, 2), (100, 4, 3, 1), (675, 3, 1, 1), (720, 5, 2, 2),
                              (1000, 2, 1, 1), (6000, 2, 1, 1), (3375, 37, 1, 1), (450, 20, 2, 1), (8100, 9, 1, 1),
                              (2187, 12, 1, 1), (3125, 8, 1, 1), (1536, 16, 1, 2), (3600, 40, 1, 1), (20, 33, 1, 1))]
    # engine 2 here = big-radix engine with the register (one butterfly per thread) kernels switched off
    cases += [(2, c) for c in ((12, 7, 2, 3), (675, 3, 1, 1), (3375, 37, 1, 1), (450, 20, 2, 1), (3600, 40, 1, 1))]
    for engine, (N, other, s1, s2) in cases:
        lib.mlb_set_option(b"mixed_registers", 0 if engine == 2 else 2)      # 2: register kernels wherever they apply
        big = (rng.standard_normal((other * s1, N * s2)) + 1j * rng.standard_normal((other * s1, N * s2))).astype(np.complex64)
        ldi = N * s2 + 3
        dbig = [torch.zeros(other * s1, ldi, dtype=torch.complex64).cuda()]
        dbig[0][:, :N * s2].copy_(torch.from_numpy(big))
        tw = torch.empty(2 * N, dtype=torch.complex64).cuda()
        _lib.check(lib.mlb_fft_twiddle(N, tw.data_ptr(), None), "tw")
        rr, rc, ro = other // 2, N // 2, (N // 2 + 1) % N
        dres = [torch.zeros(other, N + 1, dtype=torch.complex64).cuda()]
        pi_, k1 = _lib.ptr_array(dbig)
        po, k2 = _lib.ptr_array(dres)
        _lib.check(lib.mlb_fft_rows(pi_, ldi, po, N + 1, other, N, s1, s2, tw.data_ptr(), rr, rc, ro, 0, 1, None), "mixed rows")
        folded = big.astype(complex).reshape(s1, other, s2, N).sum(axis=(0, 2))
        ref = np.roll(np.fft.fft(np.roll(folded, (rr, rc), axis=(0, 1)), axis=1), ro, axis=1)
        torch.cuda.synchronize()
        assert field_error(dres[0][:, :N].cpu().numpy(), ref) < 3e-6, ("mixed rows", N)
        ldc = other + 3
        dcol = [torch.zeros(N, ldc, dtype=torch.complex64).cuda()]
        dcol[0][:, :other].copy_(torch.from_numpy(folded.T.astype(np.complex64).copy()))
        dco = [torch.zeros(N, ldc, dtype=torch.complex64).cuda()]
        pc, k3 = _lib.ptr_array(dcol)
        pco, k5 = _lib.ptr_array(dco)
        _lib.check(lib.mlb_fft_cols(pc, ldc, pco, ldc, N, other, tw.data_ptr(), ro, 1, None), "mixed cols")
        torch.cuda.synchronize()
        refc = np.roll(np.fft.fft(folded.T.astype(np.complex64).astype(complex), axis=0), ro, axis=0)
        assert field_error(dco[0][:, :other].cpu().numpy(), refc) < 3e-6, ("mixed cols", N)
        assert float(dco[0][:, other:].abs().max()) == 0.0                       # pitch padding untouched
    lib.mlb_set_option(b"mixed_registers", 1)
    # fused fold: [n_rows*s1][N*s2] input, summed over the aliased copies while loading
    n_rows, N, s1, s2 = 6, 64, 3, 4
    big = (rng.standard_normal((n_rows * s1, N * s2)) + 1j * rng.standard_normal((n_rows * s1, N * s2))).astype(np.complex64)
    dbig = [torch.from_numpy(big).cuda()]
    dres = [torch.zeros(n_rows, N, dtype=torch.complex64).cuda()]
    tw = torch.empty(2 * N, dtype=torch.complex64).cuda()
    _lib.check(lib.mlb_fft_twiddle(N, tw.data_ptr(), None), "tw")
    pi_, k1 = _lib.ptr_array(dbig)
    po, k2 = _lib.ptr_array(dres)
    _lib.check(lib.mlb_fft_rows(pi_, N * s2, po, N, n_rows, N, s1, s2, tw.data_ptr(), 2, 10, 5, 0, 1, None), "fold rows")
    torch.cuda.synchronize()
    folded = big.astype(complex).reshape(s1, n_rows, s2, N).sum(axis=(0, 2))
    ref = np.roll(np.fft.fft(np.roll(folded, (2, 10), axis=(0, 1)), axis=1), 5, axis=1)
    assert field_error(dres[0].cpu().numpy(), ref) < 3e-6
    # TMA-fed persistent kernel (256..2048 points), plain and transposed output, ragged row count
    for N, n_rows, s1, s2 in ((256, 7, 2, 3), (1024, 5, 1, 1), (2048, 3, 2, 2), (512, 301, 1, 1)):
        assert lib.mlb_fft_rows_can_transpose(N) == 1
        big = (rng.standard_normal((n_rows * s1, N * s2)) + 1j * rng.standard_normal((n_rows * s1, N * s2))).astype(np.complex64)
        dbig = [torch.from_numpy(big).cuda()]
        tw = torch.empty(2 * N, dtype=torch.complex64).cuda()
        _lib.check(lib.mlb_fft_twiddle(N, tw.data_ptr(), None), "tw")
        folded = big.astype(complex).reshape(s1, n_rows, s2, N).sum(axis=(0, 2))
        rr, rc, ro = n_rows // 2, 10, N // 2 + 3
        ref = np.roll(np.fft.fft(np.roll(folded, (rr, rc), axis=(0, 1)), axis=1), ro, axis=1)
        for tr in (0, 1):
            ldo = (n_rows + (n_rows & 1) + 2) if tr else N
            dres = [torch.zeros((N, ldo) if tr else (n_rows, ldo), dtype=torch.complex64).cuda()]
            pi_, k1 = _lib.ptr_array(dbig)
            po, k2 = _lib.ptr_array(dres)
            _lib.check(lib.mlb_fft_rows(pi_, N * s2, po, ldo, n_rows, N, s1, s2, tw.data_ptr(), rr, rc, ro, tr, 1, None), "tma rows")
            torch.cuda.synchronize()
            got = dres[0].cpu().numpy()
            got = got[:, :n_rows].T if tr else got
            assert field_error(got, ref) < 3e-6, (N, tr)
    assert lib.mlb_fft_rows_can_transpose(4096) == 0 and lib.mlb_fft_rows_can_transpose(128) == 0


def test_twiddle_float64_phase_accuracy():
    """H1: phases of ~1e4 rad must still be accurate to fp32 rounding."""
    from metalens_b200.farfield import FarfieldPlan
    M = 4096
    d = WL / 2.2
    plan = FarfieldPlan((M, 8), d, d, WL, NG, stride=(64, 1), method="dense")
    x_rel = (np.arange(M) - M // 2) * d
    ref = np.exp(-1j * (2 * np.pi * NG / WL) * np.outer(x_rel, plan.ux))
    got = plan.AxT[:, :plan.Kx].cpu().numpy()
    assert np.abs(got - ref).max() < 1.5e-7


def test_grid_violations_raise_like_reference():
    from metalens_b200.farfield import farfield_from_nearfield, farfield_from_fields
    x = np.arange(16) * 1e-7
    F = np.zeros((16, 16), complex)
    with pytest.raises(AssertionError):
        farfield_from_nearfield(F, F, F, F, x * 10, x, WL, NG)           # spacing >= lambda/2
    with pytest.raises(AssertionError):
        farfield_from_nearfield(F[:8], F, F, F, x, x, WL, NG)            # shape mismatch
    bad = x.copy(); bad[5] += 2e-9
    with pytest.raises(AssertionError):
        farfield_from_fields(F, F, F, F, bad, x, WL, NG)                 # non-uniform


def test_config2_full_size_properties():
    """BASELINE config 2 (2048^2 -> 512^2): fold and dense paths agree; a sample of bins equals the
    float64 direct sum (oracle); linearity of the aperture sums."""
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import FarfieldPlan
    M, s = 2048, 4
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 1, WL, NG)
    d = x[1] - x[0]
    dev = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    fold = FarfieldPlan((M, M), d, d, WL, NG, stride=s, method="fold")
    dense = FarfieldPlan((M, M), d, d, WL, NG, stride=s, method="dense")
    fft = FarfieldPlan((M, M), d, d, WL, NG, stride=s)
    assert fft.method == "fft"
    Pf, tf = fold.run(dev)
    Pd, td = dense.run(dev)
    Pq, tq = fft.run(dev)
    Pf, Pd, Pq = Pf.cpu().numpy(), Pd.cpu().numpy(), Pq.cpu().numpy()
    assert power_map_error(Pf, Pd) < FF_TOL
    assert power_map_error(Pq, Pd) < FF_TOL
    assert abs(tf.item() - td.item()) <= FF_TOL * abs(td.item())
    assert abs(tq.item() - td.item()) <= FF_TOL * abs(td.item())
    # oracle at a sample of bins (float64 direct sum over the whole aperture)
    ii = np.array([0, 17, 200, 255, 256, 257, 300, 511])
    jj = np.array([3, 128, 256, 260, 400])
    P_ref, _ = fo.farfield_dense(Ex, Ey, Hx, Hy, d, d, fold.ux[ii], fold.uy[jj], WL, NG)
    scale = np.nanmax(Pd)
    sub_f, sub_d, sub_q = Pf[np.ix_(ii, jj)], Pd[np.ix_(ii, jj)], Pq[np.ix_(ii, jj)]
    assert np.array_equal(np.isnan(sub_f), np.isnan(P_ref))
    fin = np.isfinite(P_ref)
    assert np.abs(sub_f - P_ref)[fin].max() / scale < FF_TOL
    assert np.abs(sub_d - P_ref)[fin].max() / scale < FF_TOL
    assert np.abs(sub_q - P_ref)[fin].max() / scale < FF_TOL
    # KAT: energy conservation of the lens aperture, total_P / P_in ~ 1
    P_in = float((Ex * np.conj(Hy) - Ey * np.conj(Hx)).real.sum()) * d * d
    assert abs(tf.item() / P_in - 1) < 5e-3
    # linearity: F(2a + b) = 2F(a) + F(b) on the aperture sums
    a1 = fold.amplitudes().clone()
    rot = [torch.from_numpy(a).cuda() for a in apertures.focusing_lens(M, 1, WL, NG, rotate=True)[:4]]
    fold.run(rot)
    a2 = fold.amplitudes().clone()
    mix = [2 * p + q for p, q in zip(dev, rot)]
    fold.run(mix)
    a3 = fold.amplitudes()
    err = (a3 - (2 * a1 + a2)).abs().max().item() / a3.abs().max().item()
    assert err < 1e-5


def test_fom_and_graph_replay():
    """A5: cone-power figure of merit, eager vs CUDA-graph replay, over a small angle sweep of tilted
    plane-wave patches (256 x 256, all bins): the peak follows sin(theta) and FOM ~ 1 inside the cone."""
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.fom import FarfieldFOM
    from oracle import farfield_oracle as fo
    M = 256
    plan = FarfieldPlan((M, M), WL / 2.2, WL / 2.2, WL, NG, stride=1)
    fom = FarfieldFOM(plan, 0.0, 0.0, 0.05)
    for deg in (5.0, 20.0, 35.0):
        Ex, Ey, Hx, Hy, x, y = apertures.tilted_te(M, WL, NG, angle_deg=deg)
        fields = [torch.from_numpy(a.astype(np.complex64)).cuda() for a in (Ex, Ey, Hx, Hy)]
        fom.set_target(np.sin(np.radians(deg)), 0.0)
        eager = fom.evaluate(fields).clone()
        replay = fom.replay().clone()                       # captures, then replays on the same inputs
        assert torch.equal(eager, replay)
        P_ref, total_ref, ux, uy, dux, duy = fo.farfield_reference_path(Ex, Ey, Hx, Hy, x, y, WL, NG)
        cone = ((ux - np.sin(np.radians(deg))) ** 2 + uy ** 2 <= 0.05 ** 2) & np.isfinite(P_ref)
        ref = np.array([P_ref[cone].sum() * dux * duy, total_ref])
        np.testing.assert_allclose(eager.cpu().numpy(), ref, rtol=2e-5)
        assert FarfieldFOM.fom(eager) > 0.95
    # replay picks up NEW inputs written into the static buffers without re-capture
    Ex, Ey, Hx, Hy, x, y = apertures.tilted_te(M, WL, NG, angle_deg=10.0)
    for i, a in enumerate((Ex, Ey, Hx, Hy)):
        fom.static_fields[i][:, :M].copy_(torch.from_numpy(a.astype(np.complex64)))
    assert FarfieldFOM.fom(fom.replay()) < 0.05             # cone still at 35 degrees -> little power there


def test_tensor_core_path_matches_reference(golden_dir):
    """tcgen05 3xTF32 complex GEMM (method 'tc'): same answer as the reference within the 1e-5
    budget, on power-of-two, ragged and arbitrary-grid cases; its complex amplitudes agree with
    the fp32 SIMT reduction to ~1e-6."""
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import FarfieldPlan, farfield_from_fields
    for name, stride in (("rand128_seed0", 1), ("lens256_seed1", 2), ("rand_45x27_seed6", 1), ("rand_48x40_seed5", (4, 2))):
        g = golden(golden_dir, name)
        Ex, Ey, Hx, Hy, x, y = CASES[name]()
        sx, sy = (stride, stride) if np.isscalar(stride) else stride
        P, total, *_ = farfield_from_fields(Ex, Ey, Hx, Hy, x, y, WL, NG, stride=stride, method="tc", p_dtype=torch.float32)
        ref = g["P"][::sx, ::sy]
        # The tensor core adds into its fp32 accumulator with truncation (round toward zero), so a COHERENT sum (the
        # focus of a lens) picks up a bias of ~2^-24 per 8-deep accumulation step.  Since round 2 the contraction is
        # split into 512-deep chunks summed with round-to-nearest (mlb_cgemm_tc_split): the bias no longer grows with
        # the aperture and the path meets north_star's 1e-5 on the lens cases too.
        assert power_map_error(P, ref) < FF_TOL, name
    Ex, Ey, Hx, Hy, x, y = apertures.gaussian_random(300, 21, WL, My=260)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (Ex, Ey, Hx, Hy)]
    ux = np.linspace(-0.7, 0.8, 150)
    uy = np.linspace(-0.6, 0.65, 131)
    tc = FarfieldPlan((300, 260), x[1] - x[0], y[1] - y[0], WL, NG, ux=ux, uy=uy, method="tc")
    simt = FarfieldPlan((300, 260), x[1] - x[0], y[1] - y[0], WL, NG, ux=ux, uy=uy, method="dense")
    P1 = tc.run(dev)[0].cpu().numpy()
    a1 = tc.amplitudes().cpu().numpy()
    P2 = simt.run(dev)[0].cpu().numpy()
    a2 = simt.amplitudes().cpu().numpy()
    assert field_error(a1, a2) < 2e-5          # truncating tensor-core accumulation, see above
    P_ref, F_ref = fo.farfield_dense(Ex, Ey, Hx, Hy, x[1] - x[0], y[1] - y[0], ux, uy, WL, NG)
    assert power_map_error(P1, P_ref) < FF_TOL and power_map_error(P2, P_ref) < FF_TOL
    assert field_error(a1[0], F_ref[0]) < 2e-5 and field_error(a2[0], F_ref[0]) < 3e-6


@pytest.mark.parametrize("M", [1024, 2048])
def test_tensor_core_path_on_large_coherent_apertures(M):
    """The case the truncating tensor-core accumulation used to fail: the focus of a large coherent lens aperture
    (depth 2 M real = M / 16 k-blocks).  With the chunked contraction the tensor path meets 1e-5 against the FFT path
    (itself pinned to the oracle on every bin, tests/test_full_configs_gpu.py)."""
    from metalens_b200.farfield import FarfieldPlan
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 17, WL, NG)
    dev = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    d = x[1] - x[0]
    ref = FarfieldPlan((M, M), d, d, WL, NG, stride=4, method="fft")
    tc = FarfieldPlan((M, M), d, d, WL, NG, stride=4, method="tc")
    P_ref, t_ref = ref.run(dev)
    P_tc, t_tc = tc.run(dev)
    assert power_map_error(P_tc.cpu().numpy(), P_ref.cpu().numpy()) < FF_TOL
    assert abs(t_tc.item() - t_ref.item()) <= FF_TOL * abs(t_ref.item())


def test_random_shapes_all_methods_against_oracle():
    """Seeded sweep over ragged aperture shapes, strides and arbitrary grids: every applicable method
    against the float64 direct sum (oracle), plus the all-zero aperture (empty input)."""
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import FarfieldPlan
    rng = np.random.default_rng(2024)
    d = WL / 2.2
    for trial in range(14):
        Mx, My = int(rng.integers(6, 150)), int(rng.integers(6, 150))
        fields = [(rng.standard_normal((Mx, My)) + 1j * rng.standard_normal((Mx, My))).astype(np.complex64) for _ in range(4)]
        dev = [torch.from_numpy(a).cuda() for a in fields]
        if trial % 2 == 0:                       # arbitrary direction-cosine grid
            ux = np.sort(rng.uniform(-0.95, 0.95, int(rng.integers(1, 70))))
            uy = np.sort(rng.uniform(-0.95, 0.95, int(rng.integers(1, 70))))
            kw, methods = dict(ux=ux, uy=uy), ("dense", "tc")
        else:                                    # strided FFT-bin grid
            sx = int(rng.choice([s for s in (1, 2, 3, 4) if Mx % s == 0 and (Mx // 2) % s == 0]))
            sy = int(rng.choice([s for s in (1, 2, 3, 4) if My % s == 0 and (My // 2) % s == 0]))
            kw, methods = dict(stride=(sx, sy)), ("auto", "dense", "tc") + (("fold",) if sx * sy > 1 else ())
        ref = None
        for m in methods:
            plan = FarfieldPlan((Mx, My), d, d, WL, NG, method=m, **kw)
            P = plan.run(dev)[0].cpu().numpy()
            if ref is None:
                ref, _ = fo.farfield_dense(*fields, d, d, plan.ux, plan.uy, WL, NG)
            if np.isfinite(ref).any():
                assert power_map_error(P, ref) < (3e-5 if m == "tc" else FF_TOL), (trial, Mx, My, m, kw.get("stride"))
            else:
                assert np.isnan(P).all()
    zero = [torch.zeros(64, 64, dtype=torch.complex64).cuda() for _ in range(4)]
    plan = FarfieldPlan((64, 64), d, d, WL, NG, stride=1)
    P, total = plan.run(zero)
    P = P.cpu().numpy()
    assert total.item() == 0.0 and np.all((P == 0) | np.isnan(P)) and np.isnan(P).sum() > 0


def test_config3_full_size_fft_path():
    """BASELINE config 3 size (4096^2 -> 1024^2): the TMA-fed fold+FFT path against the float64 direct
    sum on a sample of bins, and against the stand-alone fold + tiled reduction."""
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import FarfieldPlan
    M, s = 4096, 4
    wl, ng = 635e-9, apertures.N_GLASS[635]
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 7, wl, ng)
    d = x[1] - x[0]
    dev = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    fft = FarfieldPlan((M, M), d, d, wl, ng, stride=s)
    assert fft.method == "fft"
    Pq, tq = fft.run(dev)
    Pq = Pq.cpu().numpy()
    fold = FarfieldPlan((M, M), d, d, wl, ng, stride=s, method="fold")
    Pf, tf = fold.run(dev)
    assert power_map_error(Pq, Pf.cpu().numpy()) < FF_TOL
    assert abs(tq.item() - tf.item()) <= FF_TOL * abs(tf.item())
    ii = np.array([0, 100, 511, 512, 513, 700, 1023])
    jj = np.array([5, 512, 520, 900])
    P_ref, _ = fo.farfield_dense(Ex, Ey, Hx, Hy, d, d, fft.ux[ii], fft.uy[jj], wl, ng)
    sub = Pq[np.ix_(ii, jj)]
    assert np.array_equal(np.isnan(sub), np.isnan(P_ref))
    fin = np.isfinite(P_ref)
    assert np.abs(sub - P_ref)[fin].max() / np.nanmax(Pq) < FF_TOL
    P_in = float((Ex * np.conj(Hy) - Ey * np.conj(Hx)).real.sum()) * d * d
    assert abs(tq.item() / P_in - 1) < 5e-3


def test_long_column_transform_all_bins():
    """4096 x 256 aperture, all FFT bins: the x transform has 4096 points and takes the four-step column
    pass (its input buffer is scratch); compared with the float64 direct sum on a sample of bins and
    with the tiled reduction; a second run() must give the same answer (scratch is re-filled)."""
    from oracle import farfield_oracle as fo
    from metalens_b200.farfield import FarfieldPlan
    Mx, My = 4096, 256
    Ex, Ey, Hx, Hy, x, y = apertures.gaussian_random(Mx, 33, WL, My=My)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (Ex, Ey, Hx, Hy)]
    plan = FarfieldPlan((Mx, My), x[1] - x[0], y[1] - y[0], WL, NG, stride=1)
    assert plan.method == "fft"
    P1 = plan.run(dev)[0].clone()
    P2 = plan.run(dev)[0].clone()
    assert bool(((P1 == P2) | (torch.isnan(P1) & torch.isnan(P2))).all())
    ii = np.array([0, 1, 63, 64, 65, 2047, 2048, 2049, 4032, 4095])
    jj = np.array([0, 17, 128, 200, 255])
    P_ref, _ = fo.farfield_dense(Ex, Ey, Hx, Hy, x[1] - x[0], y[1] - y[0], plan.ux[ii], plan.uy[jj], WL, NG)
    sub = P1.cpu().numpy()[np.ix_(ii, jj)]
    assert np.array_equal(np.isnan(sub), np.isnan(P_ref))
    fin = np.isfinite(P_ref)
    assert np.abs(sub - P_ref)[fin].max() / np.nanmax(P1.cpu().numpy()) < FF_TOL


def _same(a, b):
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())


@pytest.mark.parametrize("wide", [0, 1, 16])
@pytest.mark.parametrize("shape,stride", [((256, 256), 1), ((1024, 1024), 4), ((1024, 2048), (4, 4)),
                                          ((2048, 512), (2, 2)), ((2048, 1024), (1, 4)), ((512, 300), (1, 1)),
                                          ((4096, 300), (1, 1)), ((8192, 150), (1, 1))])
def test_fused_cols_power_matches_separate_kernels(shape, stride, wide):
    """Column pass + radiated power in one kernel (aperture sums never stored) against the separate column
    pass + epilogue on the same plan geometry: same NaN mask, P and total_P to fp32 rounding; column lengths
    256..2048 with the radix-4 kernel (fft_cols_power_kernel, both tile widths) and 256..8192 with the
    radix-16 engine (fft16_cols_power_kernel, wide == 16); square and rectangular, ragged column counts."""
    from metalens_b200 import _lib
    from metalens_b200.farfield import FarfieldPlan
    lib = _lib.load()
    Mx, My = shape
    if wide != 16 and Mx // (stride if np.isscalar(stride) else stride[0]) > 2048:
        pytest.skip("the radix-4 fused kernel stops at 2048 points")
    Ex, Ey, Hx, Hy, x, y = apertures.gaussian_random(Mx, 71, WL, My=My)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (Ex, Ey, Hx, Hy)]
    if wide == 16:
        lib.mlb_set_option(b"cols_engine", 1)
        lib.mlb_set_option(b"rows_engine", 2)
        lib.mlb_set_option(b"r16_min_lg", 8)
        if Mx >= 4096:
            lib.mlb_set_option(b"cols_strip_mb", 1 if Mx == 4096 else 0)     # 4096: ten 32-column strips (ragged last)
    else:
        lib.mlb_set_option(b"cols_engine", 0)
        lib.mlb_set_option(b"rows_engine", 0)
        lib.mlb_set_option(b"cols_power_wide", wide)
    try:
        fused = FarfieldPlan((Mx, My), x[1] - x[0], y[1] - y[0], WL, NG, stride=stride, fuse_power="always")
        plain = FarfieldPlan((Mx, My), x[1] - x[0], y[1] - y[0], WL, NG, stride=stride, fuse_power=False)
        assert fused.method == "fft" and fused.fused and not plain.fused
        assert [s[0] for s in fused.steps(dev)][-1] == "fft_cols_power"
        P1, t1 = fused.run(dev)
        P2, t2 = plain.run(dev)
        P1, P2 = P1.cpu().numpy(), P2.cpu().numpy()
        assert np.isnan(P1).sum() > 0 or max(fused.ux.max(), fused.uy.max()) < 0.7
        assert power_map_error(P1, P2) < 2e-6
        assert abs(t1.item() - t2.item()) <= 2e-6 * abs(t2.item())
        # the aperture sums are still available on request (re-runs the separate passes)
        a1, a2 = fused.amplitudes(), plain.amplitudes()
        assert torch.equal(a1, a2)
        # incoherent accumulation through the fused kernel
        Pa, ta = fused.run_incoherent([dev, dev])
        Pa = Pa.cpu().numpy()
        fin = np.isfinite(P2)
        assert np.array_equal(np.isnan(Pa), np.isnan(P2))
        assert np.abs(Pa[fin] - 2 * P2[fin]).max() <= 4e-6 * P2[fin].max()
        assert abs(ta.item() - 2 * t2.item()) <= 4e-6 * abs(t2.item())
    finally:
        lib.mlb_set_option(b"cols_power_wide", -1)
        lib.mlb_set_option(b"cols_engine", 1)
        lib.mlb_set_option(b"rows_engine", 2)
        lib.mlb_set_option(b"cols_strip_mb", 0)
        lib.mlb_set_option(b"r16_min_lg", 10)


@pytest.mark.parametrize("name,stride", [("lens256_seed1", 1), ("lens256_seed1_rot", 1)])
def test_fused_path_against_reference_golden(name, stride, golden_dir):
    """The fused float32 path against the unmodified reference's output (committed fixture)."""
    from metalens_b200.farfield import FarfieldPlan
    g = golden(golden_dir, name)
    Ex, Ey, Hx, Hy, x, y = CASES[name]()
    plan = FarfieldPlan(Ex.shape, x[1] - x[0], y[1] - y[0], WL, NG, stride=stride)
    assert plan.fused
    P, total = plan.run([torch.from_numpy(np.ascontiguousarray(a.astype(np.complex64))).cuda() for a in (Ex, Ey, Hx, Hy)])
    assert power_map_error(P.cpu().numpy(), g["P"]) < FF_TOL
    assert abs(total.item() - g["total_P"]) <= FF_TOL * abs(g["total_P"])


def test_options_api():
    from metalens_b200 import _lib
    lib = _lib.load()
    assert lib.mlb_get_option(b"rows_l2_evict_first") in (0, 1)
    assert lib.mlb_get_option(b"no_such_option") == -1
    assert lib.mlb_set_option(b"no_such_option", 1) != 0
    assert b"unknown option" in lib.mlb_last_error()
    assert lib.mlb_set_option(b"rows_ctas_per_sm", 7) != 0
    assert lib.mlb_fft_cols_power_blocks(1000, 64) == 0 and lib.mlb_fft_cols_power_blocks(16384, 64) == 0
    assert lib.mlb_get_option(b"cols_engine") == 1 and lib.mlb_get_option(b"rows_engine") == 2
    assert lib.mlb_get_option(b"r16_min_lg") == 10 and lib.mlb_fft_cols_power_blocks(512, 512) == 128     # radix-4 tile
    assert lib.mlb_get_option(b"cols_strip_mb") == 0 and lib.mlb_set_option(b"cols_strip_mb", -1) != 0
    assert lib.mlb_fft_cols_power_blocks(1024, 1023) == 256 and lib.mlb_fft_cols_power_blocks(8192, 64) == 16 * 8
    lib.mlb_set_option(b"cols_engine", 0)
    try:
        assert lib.mlb_fft_cols_power_blocks(4096, 64) == 0
        lib.mlb_set_option(b"cols_power_wide", 0)
        assert lib.mlb_fft_cols_power_blocks(1024, 1023) == 512
        lib.mlb_set_option(b"cols_power_wide", -1)
        assert lib.mlb_fft_cols_power_blocks(1024, 1023) == 256 and lib.mlb_fft_cols_power_blocks(512, 512) == 128
    finally:
        lib.mlb_set_option(b"cols_engine", 1)


@pytest.mark.parametrize("evict,per_sm", [(0, 0), (1, 1), (1, 2)])
def test_row_pass_options_do_not_change_results(evict, per_sm):
    """L2 evict-first streaming and the resident-CTA count of the TMA-fed row pass are pure scheduling knobs."""
    from metalens_b200 import _lib
    from metalens_b200.farfield import FarfieldPlan
    lib = _lib.load()
    M = 1024
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 5, WL, NG)
    dev = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    plan = FarfieldPlan((M, M), x[1] - x[0], x[1] - x[0], WL, NG, stride=4)
    ref = plan.run(dev)[0].clone()
    lib.mlb_set_option(b"rows_l2_evict_first", evict)
    lib.mlb_set_option(b"rows_ctas_per_sm", per_sm)
    try:
        assert _same(plan.run(dev)[0], ref)
    finally:
        lib.mlb_set_option(b"rows_l2_evict_first", 1)
        lib.mlb_set_option(b"rows_ctas_per_sm", 0)


@pytest.mark.parametrize("M,stride,method", [(1024, 4, "auto"), (512, 2, "auto"), (256, 4, "fold"), (96, 1, "dense")])
def test_pipelined_tiles_equal_sequential(M, stride, method):
    """ShardedFarfield.run(overlap=True) pipelines the local tiles over two streams (aperture pass of tile k+1
    over the column pass / epilogue of tile k); results must equal the one-stream run bit for bit, over
    several back-to-back steps (buffer reuse across steps), and so must a captured CUDA graph of the step."""
    from metalens_b200.farfield import FarfieldPlan
    from metalens_b200.sharding import ShardedFarfield
    n_items = 3
    K = M // stride
    fields, d = {}, None
    for i in range(n_items):
        Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 90 + i, WL, NG, rotate=bool(i % 2))
        d = x[1] - x[0]
        fields[i] = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]

    def make_plan(item, r0, r1):
        return FarfieldPlan((M, M), d, d, WL, NG, stride=stride, method=method)
    sh = ShardedFarfield(n_items, K, make_plan, rank=0, world=1)
    P_seq, tot_seq = sh.run(lambda i: fields[i])
    torch.cuda.synchronize()
    P_seq = P_seq.clone()
    tot_seq = [t.clone() for t in tot_seq]
    for _ in range(4):
        P_ov, tot_ov = sh.run(lambda i: fields[i], overlap=True)
    sh.finish()
    torch.cuda.synchronize()
    assert _same(P_ov, P_seq)
    assert all(torch.equal(a, b) for a, b in zip(tot_ov, tot_seq))
    # one captured graph of the whole step; new apertures are written into the same buffers
    sh.capture(lambda i: fields[i])
    for i in range(n_items):
        for f in fields[i]:
            f.mul_(2.0)
    P_g, tot_g = sh.replay()
    torch.cuda.synchronize()
    fin = torch.isfinite(P_seq)
    assert _same(torch.isnan(P_g), torch.isnan(P_seq))
    assert float((P_g[fin] - 4 * P_seq[fin]).abs().max()) <= 1e-6 * float(P_seq[fin].max()) * 4
    for _ in range(3):
        P_g2, _t = sh.replay()
    torch.cuda.synchronize()
    assert _same(P_g2, P_g)


def test_radix16_row_kernels_match_numpy():
    """fft16_rows_kernel (register-resident radix-16 butterflies, csrc/fft16.cuh) against numpy.fft for every
    supported length 256..8192: fftshift rolls, ragged row counts (several rows per CTA), aperture fold."""
    from metalens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(16)
    lib.mlb_set_option(b"rows_engine", 1)
    lib.mlb_set_option(b"r16_min_lg", 8)
    try:
        for N, n_rows, s1, s2 in ((256, 37, 1, 1), (512, 9, 1, 1), (1024, 5, 1, 1), (2048, 3, 1, 1), (4096, 3, 1, 1),
                                  (8192, 2, 1, 1), (256, 7, 2, 3), (1024, 6, 4, 4), (4096, 2, 2, 1), (8192, 1, 1, 2)):
            big = (rng.standard_normal((n_rows * s1, N * s2)) + 1j * rng.standard_normal((n_rows * s1, N * s2))).astype(np.complex64)
            ldi = N * s2 + 2
            dbig = [torch.zeros(n_rows * s1, ldi, dtype=torch.complex64).cuda() for _ in range(2)]
            dbig[0][:, :N * s2].copy_(torch.from_numpy(big))
            dbig[1][:, :N * s2].copy_(torch.from_numpy(2 * big))
            tw = torch.empty(2 * N, dtype=torch.complex64).cuda()
            _lib.check(lib.mlb_fft_twiddle(N, tw.data_ptr(), None), "tw")
            folded = big.astype(complex).reshape(s1, n_rows, s2, N).sum(axis=(0, 2))
            rr, rc, ro = n_rows // 2, (N // 2 + 5) % N, N // 2 + 3
            ref = np.roll(np.fft.fft(np.roll(folded, (rr, rc), axis=(0, 1)), axis=1), ro, axis=1)
            dres = [torch.zeros(n_rows, N + 4, dtype=torch.complex64).cuda() for _ in range(2)]
            pi_, k1 = _lib.ptr_array(dbig)
            po, k2 = _lib.ptr_array(dres)
            _lib.check(lib.mlb_fft_rows(pi_, ldi, po, N + 4, n_rows, N, s1, s2, tw.data_ptr(), rr, rc, ro, 0, 2, None), "r16 rows")
            torch.cuda.synchronize()
            assert field_error(dres[0][:, :N].cpu().numpy(), ref) < 3e-6, (N, s1, s2)
            assert field_error(dres[1][:, :N].cpu().numpy(), 2 * ref) < 3e-6, (N, s1, s2)
            assert float(dres[0][:, N:].abs().max()) == 0.0                     # pitch padding untouched
    finally:
        lib.mlb_set_option(b"rows_engine", 2)
        lib.mlb_set_option(b"r16_min_lg", 10)


def test_radix16_column_kernels_match_numpy():
    """fft16_cols_kernel (direct, 256..2048 points) and the two-pass 16 x B decomposition for 4096 / 8192 points
    (fft16_cols_first_kernel in place on the input + fft16_cols_kernel) against numpy.fft: output roll, ragged
    column counts (partial column tiles), two fields per launch."""
    from metalens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(17)
    lib.mlb_set_option(b"cols_engine", 1)
    lib.mlb_set_option(b"r16_min_lg", 8)
    try:
        for N, n_cols, strip_mb in ((256, 45, 0), (512, 17, 0), (1024, 9, 0), (2048, 6, 0), (4096, 37, 0),
                                    (8192, 33, 0), (1024, 64, 0), (4096, 77, 1), (8192, 70, 48), (8192, 70, 2)):
            lib.mlb_set_option(b"cols_strip_mb", strip_mb)     # 1-2 MB: several 32-column strips with a ragged tail
            a = [(rng.standard_normal((N, n_cols)) + 1j * rng.standard_normal((N, n_cols))).astype(np.complex64) for _ in range(2)]
            tw = torch.empty(2 * N, dtype=torch.complex64).cuda()
            _lib.check(lib.mlb_fft_twiddle(N, tw.data_ptr(), None), "tw")
            ldc = n_cols + (n_cols & 1) + 2
            din = [torch.zeros(N, ldc, dtype=torch.complex64).cuda() for _ in a]
            for d, v in zip(din, a):
                d[:, :n_cols].copy_(torch.from_numpy(v))
            dout = [torch.zeros(N, ldc, dtype=torch.complex64).cuda() for _ in a]
            ro = (N // 2 + 1) % N
            pc, k3 = _lib.ptr_array(din)
            po, k4 = _lib.ptr_array(dout)
            if N >= 4096:
                assert lib.mlb_fft_cols(pc, ldc, pc, ldc, N, n_cols, tw.data_ptr(), ro, 2, None) != 0   # needs out != in
            _lib.check(lib.mlb_fft_cols(pc, ldc, po, ldc, N, n_cols, tw.data_ptr(), ro, 2, None), "r16 cols")
            torch.cuda.synchronize()
            for v, d in zip(a, dout):
                ref = np.roll(np.fft.fft(v.astype(complex), axis=0), ro, axis=0)
                assert field_error(d[:, :n_cols].cpu().numpy(), ref) < 3e-6, ("cols", N)
                assert float(d[:, n_cols:].abs().max()) == 0.0
            if N <= 2048:                                                   # in place is allowed for the direct pass
                _lib.check(lib.mlb_fft_cols(pc, ldc, pc, ldc, N, n_cols, tw.data_ptr(), ro, 2, None), "r16 cols in place")
                torch.cuda.synchronize()
                assert torch.equal(din[0], dout[0]) and torch.equal(din[1], dout[1])
    finally:
        lib.mlb_set_option(b"cols_engine", 1)
        lib.mlb_set_option(b"cols_strip_mb", 0)
        lib.mlb_set_option(b"r16_min_lg", 10)


@pytest.mark.parametrize("M", [3375, 2700])
def test_reference_default_grid_full_size_properties(M):
    """The reference's own usage at full size: a good_fft_number() aperture (3375 = 3^3 5^3, 2700 = 2^2 3^3 5^2),
    ALL FFT bins, through the big-radix mixed engine.  Size-independent properties: a sample of bins against the
    float64 direct sum (oracle), Parseval on the complex aperture sums, linearity, energy conservation of the lens
    (KAT), and agreement of the register kernels with the shared-memory ones."""
    from oracle import farfield_oracle as fo
    from metalens_b200 import _lib
    from metalens_b200.farfield import FarfieldPlan
    lib = _lib.load()
    Ex, Ey, Hx, Hy, x, y = apertures.focusing_lens(M, 4, WL, NG)
    d = x[1] - x[0]
    dev = [torch.from_numpy(a).cuda() for a in (Ex, Ey, Hx, Hy)]
    plan = FarfieldPlan((M, M), d, d, WL, NG, stride=1)
    assert plan.method == "fft"
    P, total = plan.run(dev)
    P = P.clone()
    A = plan.amplitudes()                                           # (4, M, M) complex aperture sums
    # Parseval: sum |F|^2 = M^2 sum |J|^2 per field
    for f in range(4):
        e_in = float((dev[f].abs().double() ** 2).sum())
        e_out = float((A[f].abs().double() ** 2).sum())
        assert abs(e_out - M * M * e_in) <= 2e-5 * max(M * M * e_in, 1e-300)
    # oracle at a sample of bins
    ii = np.array([0, 1, M // 3, M // 2 - 1, M // 2, M // 2 + 1, M - 7, M - 1])
    jj = np.array([2, M // 5, M // 2, M // 2 + 9, M - 1])
    P_ref, F_ref = fo.farfield_dense(Ex, Ey, Hx, Hy, d, d, plan.ux[ii], plan.uy[jj], WL, NG)
    Pn = P.cpu().numpy()
    sub = Pn[np.ix_(ii, jj)]
    assert np.array_equal(np.isnan(sub), np.isnan(P_ref))
    fin = np.isfinite(P_ref)
    assert np.abs(sub - P_ref)[fin].max() / np.nanmax(Pn) < FF_TOL
    for f in (0, 3):
        got = A[f].cpu().numpy()[np.ix_(ii, jj)]
        assert np.abs(got - F_ref[f]).max() <= 2e-5 * A[f].abs().max().item()
    P_in = float((Ex * np.conj(Hy) - Ey * np.conj(Hx)).real.sum()) * d * d
    assert abs(total.item() / P_in - 1) < 5e-3
    # linearity of the aperture sums: FFT(2 a + i b) = 2 FFT(a) + i FFT(b)
    other = [torch.roll(t, shifts=(5, -3), dims=(0, 1)).contiguous() for t in dev]
    combo = [2.0 * a + 1j * b for a, b in zip(dev, other)]
    A1 = A.clone()
    plan.run(other); A2 = plan.amplitudes().clone()
    plan.run(combo); A3 = plan.amplitudes()
    assert ((A3 - (2.0 * A1 + 1j * A2)).abs().max() / A3.abs().max()).item() < 2e-5
    # shared-memory (loop) kernels of the same engine give the same map
    lib.mlb_set_option(b"mixed_registers", 0)
    try:
        P0, t0 = plan.run(dev)
        assert power_map_error(P0.cpu().numpy(), Pn) < 1e-6 and abs(t0.item() - total.item()) <= 1e-6 * abs(total.item())
    finally:
        lib.mlb_set_option(b"mixed_registers", 1)
