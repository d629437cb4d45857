"""GPU: the compile-time plans of the big-radix mixed engine (csrc/fftmix.cuh, mix2_ct_kernel) -- every plan the
dispatch serves with them, against numpy's FFT of the same complex64 data (what the reference's caller runs,
nearfield_farfield.py:18-20) and against the generic run-time-plan kernels they replace."""
import numpy as np
import pytest

from parity import field_error

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

# rows: the three-stage lengths 2048 < N <= 4096 of the form 2^a 3^b 5^c whose busiest stage fills a 256-thread CTA
ROW_LENGTHS = (2160, 2250, 2304, 2400, 2560, 2700, 2880, 3072, 3375, 3600, 3840)
# second column pass / short direct columns: two-stage sub-lengths with 13..16 butterflies per stage and column
COL_SUBLENGTHS = (30, 45, 60, 75, 90, 120, 135, 144, 150, 160, 180, 192, 225, 240)


def _twiddle(lib, _lib, N):
    tw = torch.empty(2 * N, dtype=torch.complex64).cuda()
    _lib.check(lib.mlb_fft_twiddle(N, tw.data_ptr(), None), "tw")
    return tw


def _rows(lib, _lib, big, N, other, s1, s2, rolls):
    rr, rc, ro = rolls
    ldi = N * s2 + 3
    din = [torch.zeros(other * s1, ldi, dtype=torch.complex64).cuda()]
    din[0][:, :N * s2].copy_(torch.from_numpy(big))
    dout = [torch.zeros(other, N + 1, dtype=torch.complex64).cuda()]
    pi_, k1 = _lib.ptr_array(din)
    po, k2 = _lib.ptr_array(dout)
    tw = _twiddle(lib, _lib, N)
    _lib.check(lib.mlb_fft_rows(pi_, ldi, po, N + 1, other, N, s1, s2, tw.data_ptr(), rr, rc, ro, 0, 1, None), "rows")
    torch.cuda.synchronize()
    assert float(dout[0][:, N:].abs().max()) == 0.0                              # pitch padding untouched
    return dout[0][:, :N].cpu().numpy()


def _cols(lib, _lib, data, N, n_cols, ro):
    ldc = n_cols + 3
    din = [torch.zeros(N, ldc, dtype=torch.complex64).cuda() for _ in range(2)]
    for d in din:
        d[:, :n_cols].copy_(torch.from_numpy(data))
    dout = [torch.zeros(N, ldc, dtype=torch.complex64).cuda() for _ in range(2)]
    pc, k1 = _lib.ptr_array(din)
    po, k2 = _lib.ptr_array(dout)
    tw = _twiddle(lib, _lib, N)
    _lib.check(lib.mlb_fft_cols(pc, ldc, po, ldc, N, n_cols, tw.data_ptr(), ro, 2, None), "cols")
    torch.cuda.synchronize()
    assert float(dout[0][:, n_cols:].abs().max()) == 0.0
    assert torch.equal(dout[0], dout[1])                                          # both batch entries, same data
    return dout[0][:, :n_cols].cpu().numpy()


def test_option_is_listed():
    from metalens_b200 import _lib
    lib = _lib.load()
    assert lib.mlb_get_option(b"mixed_compiled") == 1
    assert lib.mlb_set_option(b"mixed_compiled", 0) == 0 and lib.mlb_get_option(b"mixed_compiled") == 0
    assert lib.mlb_set_option(b"mixed_compiled", 1) == 0


@pytest.mark.parametrize("N", ROW_LENGTHS)
def test_compiled_row_plans(N):
    from metalens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(N)
    try:
        for other, s1, s2, occ in ((5, 1, 1, 0), (3, 2, 2, 3), (1, 1, 1, 4)):
            big = (rng.standard_normal((other * s1, N * s2)) + 1j * rng.standard_normal((other * s1, N * s2))).astype(np.complex64)
            rolls = (other // 2, N // 2, (N // 2 + 1) % N)
            folded = big.astype(complex).reshape(s1, other, s2, N).sum(axis=(0, 2))
            ref = np.roll(np.fft.fft(np.roll(folded, rolls[:2], axis=(0, 1)), axis=1), rolls[2], axis=1)
            lib.mlb_set_option(b"mixed_occupancy", occ)
            out = {}
            for compiled in (1, 0):
                lib.mlb_set_option(b"mixed_compiled", compiled)
                out[compiled] = _rows(lib, _lib, big, N, other, s1, s2, rolls)
                assert field_error(out[compiled], ref) < 3e-6, ("rows", N, compiled, other, s1, s2)
            assert field_error(out[1], out[0]) < 1e-6, ("rows, compiled vs generic", N)
    finally:
        lib.mlb_set_option(b"mixed_compiled", 1)
        lib.mlb_set_option(b"mixed_occupancy", 0)


@pytest.mark.parametrize("B", COL_SUBLENGTHS)
def test_compiled_column_plans(B):
    """Direct columns of length B, and N = 16 B / 15 B columns in two passes whose second pass runs the plan of B."""
    from metalens_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(B)
    lengths = [B] + [A * B for A in (16, 15) if A * B <= 8192 and A * B >= 271 and (A * B) & (A * B - 1)]
    try:
        for N in lengths:
            for n_cols, occ in ((37, 0), (9, 3)):
                data = (rng.standard_normal((N, n_cols)) + 1j * rng.standard_normal((N, n_cols))).astype(np.complex64)
                ro = (N // 2 + 1) % N
                ref = np.roll(np.fft.fft(data.astype(complex), axis=0), ro, axis=0)
                lib.mlb_set_option(b"mixed_occupancy", occ)
                out = {}
                for compiled in (1, 0):
                    lib.mlb_set_option(b"mixed_compiled", compiled)
                    out[compiled] = _cols(lib, _lib, data, N, n_cols, ro)
                    assert field_error(out[compiled], ref) < 3e-6, ("cols", N, B, compiled, n_cols)
                assert field_error(out[1], out[0]) < 1e-6, ("cols, compiled vs generic", N, B)
    finally:
        lib.mlb_set_option(b"mixed_compiled", 1)
        lib.mlb_set_option(b"mixed_occupancy", 0)
