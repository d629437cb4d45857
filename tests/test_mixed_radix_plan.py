"""CPU: numpy emulation of the big-radix mixed FFT engine (csrc/fftmix.cuh) -- the radix plan of mix2_plan(), the
Stockham stage index math, the composite in-register butterflies with their output slot map, the multiply-high
index reciprocals and the two-pass (N = A x B) column decomposition -- against numpy.fft for 5-smooth lengths
(what good_fft_number() yields, nearfield.py:30-36)."""
import numpy as np
import pytest

SPLIT = {6: (3, 2), 9: (3, 3), 10: (5, 2), 12: (4, 3), 15: (5, 3)}


def plan(N):
    """mix2_plan(): odd radices first, then the even ones."""
    a = b = c = 0
    n = N
    while n % 2 == 0: a += 1; n //= 2
    while n % 3 == 0: b += 1; n //= 3
    while n % 5 == 0: c += 1; n //= 5
    assert n == 1 and N >= 2
    r = []
    while b >= 1 and c >= 1: r.append(15); b -= 1; c -= 1
    while b >= 2: r.append(9); b -= 2
    tens = 0
    while c >= 1:
        if tens < a: tens += 1
        else: r.append(5)
        c -= 1
    a -= tens
    if b == 1:
        if a >= 2: r.append(12); a -= 2
        elif a >= 1: r.append(6); a -= 1
        else: r.append(3)
    r += [10] * tens
    while a >= 4: r.append(16); a -= 4
    if a: r.append(1 << a)
    assert int(np.prod(r)) == N
    return r


def dft_mix(v):
    """in-register DFT of len(v) points; returns (slots, slot_of_output)"""
    R = len(v)
    if R not in SPLIT:
        return np.fft.fft(v), list(range(R))
    A, B = SPLIT[R]
    v = np.array(v, dtype=complex)
    for n2 in range(B):
        t = np.fft.fft(v[[n2 + B * n1 for n1 in range(A)]])
        for k1 in range(A):
            v[n2 + B * k1] = t[k1] * np.exp(-2j * np.pi * ((n2 * k1) % R) / R)
    for k1 in range(A):
        t = np.fft.fft(v[[n2 + B * k1 for n2 in range(B)]])
        for k2 in range(B):
            v[k2 + B * k1] = t[k2]
    return v, [(q // A) + B * (q % A) for q in range(R)]


def magic(d):
    return 0 if d <= 1 else ((1 << 32) + d - 1) // d


def fft_mix(x, tw, tw_mul):
    """all stages of one transform; tw = plain table of the (possibly longer) full length"""
    N = len(x)
    cur = np.array(x, dtype=complex)
    Ns = 1
    for R in plan(N):
        per = N // R
        y = np.empty(N, dtype=complex)
        twstep = (N // (Ns * R)) * tw_mul
        m = magic(Ns)
        for j in range(per):
            k = 0 if Ns == 1 else j - ((j * m) >> 32) * Ns
            assert k == j % Ns
            v = cur[[j + r * per for r in range(R)]]
            if Ns > 1:
                w1 = tw[k * twstep]
                pw = [1.0, w1]
                for r in range(2, R):
                    pw.append(pw[r >> 1] * pw[r - (r >> 1)])
                v = v * np.array(pw[:R])
            V, slot = dft_mix(v)
            for q in range(R):
                y[(j - k) * R + k + q * Ns] = V[slot[q]]
        cur = y
        Ns *= R
    return cur


@pytest.mark.parametrize("N", [6, 12, 45, 100, 225, 450, 675, 720, 1000, 1536, 2187, 3125, 3375])
def test_direct_transform(N):
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    assert np.abs(fft_mix(x, tw, 1) - np.fft.fft(x)).max() < 1e-9 * N


@pytest.mark.parametrize("N,A", [(3375, 15), (3600, 16), (450, 15), (8100, 15), (2187, 9), (3125, 5), (1536, 16), (1350, 15), (20, 10)])
def test_two_pass_columns(N, A):
    """pass 1: A-point DFTs over rows n2 + B n1, times W_N^(n2 k1), stored at row k1 B + n2; pass 2: B-point
    transforms of the contiguous rows of each k1 (table stride A), output k2 at row k1 + A k2."""
    B = N // A
    rng = np.random.default_rng(N + A)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    d = x.copy()
    for n2 in range(B):
        V, slot = dft_mix(d[[n2 + m * B for m in range(A)]])
        w1 = tw[n2]
        pw = [1.0, w1]
        for r in range(2, A):
            pw.append(pw[r >> 1] * pw[r - (r >> 1)])
        for k1 in range(A):
            d[k1 * B + n2] = V[slot[k1]] * (pw[k1] if k1 else 1.0)
    out = np.empty(N, dtype=complex)
    for g in range(A):
        sub = fft_mix(d[g * B:(g + 1) * B], tw, A)
        for k2 in range(B):
            out[g + A * k2] = sub[k2]
    assert np.abs(out - np.fft.fft(x)).max() < 1e-9 * N


def test_reciprocals_are_exact():
    """floor(n / d) by multiply-high with ceil(2^32 / d) for every index the kernels form (n < 65536, d <= 8192)."""
    n = np.arange(65536, dtype=np.uint64)
    for d in list(range(2, 70)) + [225, 240, 405, 1125, 3375, 4096, 4050, 8100, 8192]:
        m = np.uint64(magic(d))
        lim = n[n * np.uint64(d) < (1 << 27)] if d > 64 else n
        assert np.array_equal((lim * m) >> np.uint64(32), lim // np.uint64(d)), d


def test_library_planner_matches_the_emulated_one():
    """mlb_fft_mixed_plan (host only) returns exactly the radix sequence emulated above, for every 5-smooth length
    the engine serves, with odd radices first and the pad rule that follows from the first radix."""
    import ctypes
    from metalens_b200 import _lib
    lib = _lib.load()
    sizes = sorted({2 ** a * 3 ** b * 5 ** c for a in range(14) for b in range(9) for c in range(6)
                    if 2 <= 2 ** a * 3 ** b * 5 ** c <= 8192})
    assert len(sizes) > 150
    for N in sizes:
        r = (ctypes.c_int * 8)()
        sh = ctypes.c_int()
        ns = lib.mlb_fft_mixed_plan(N, r, ctypes.byref(sh))
        want = plan(N)
        assert ns == len(want) and list(r)[:ns] == want and all(v == 0 for v in list(r)[ns:]), N
        assert sh.value == (30 if want[0] % 2 else 4)
        odd = [v % 2 for v in want]
        assert odd == sorted(odd, reverse=True), (N, want)         # odd radices first
    assert lib.mlb_fft_mixed_plan(14, (ctypes.c_int * 8)(), None) < 0   # 2 * 7: not 5-smooth


def test_every_register_kernel_plan_is_compiled():
    """The dispatch of mlb_fft_rows / mlb_fft_cols (csrc/fft.cu) sends a transform to the one-butterfly-per-thread
    register kernels when its busiest stage fills >= 80 % of a 256-thread CTA: rows of N >= 2048 points (one row per
    CTA), column (sub-)transforms with 16 columns per CTA.  Every such plan must have its compile-time instantiation
    (mix2_ct_kernel, fftmix.cuh) -- the generic run-time-plan kernel spills at 128 registers -- and no other plan may."""
    from metalens_b200 import _lib
    lib = _lib.load()
    sizes = sorted({2 ** a * 3 ** b * 5 ** c for a in range(14) for b in range(9) for c in range(6)
                    if 2 <= 2 ** a * 3 ** b * 5 ** c <= 8192})
    rows, cols = [], []
    for N in sizes:
        if N & (N - 1) == 0:
            assert lib.mlb_fft_mixed_compiled(N, 0) == 0 and lib.mlb_fft_mixed_compiled(N, 1) == 0   # radix-16 engine's lengths
            continue
        busiest = max(N // r for r in plan(N))
        # rows (fft_rows_impl): rl = 256 // busiest rows per CTA, register variant if N >= 2048 and rl * busiest >= 205
        rl = min(256 // busiest, 16) if busiest <= 256 else 0
        want_row = rl >= 1 and N >= 2048 and rl * busiest >= 205
        assert not want_row or rl == 1
        assert lib.mlb_fft_mixed_compiled(N, 0) == (1 if want_row else 0), ("rows", N, plan(N))
        rows += [N] if want_row else []
        # columns (mlb_fft_cols, direct or second pass): rl = largest power of two <= 32 with rl * busiest <= 256
        rl = 1
        while rl < 32 and rl * 2 * busiest <= 256:
            rl *= 2
        want_col = rl * busiest <= 256 and rl >= 8 and rl * busiest >= 205
        assert not want_col or rl == 16
        assert lib.mlb_fft_mixed_compiled(N, 1) == (1 if want_col else 0), ("cols", N, plan(N))
        cols += [N] if want_col else []
    assert rows == [2160, 2250, 2304, 2400, 2560, 2700, 2880, 3072, 3375, 3600, 3840]
    assert cols == [30, 45, 60, 75, 90, 120, 135, 144, 150, 160, 180, 192, 225, 240]
    assert lib.mlb_fft_mixed_compiled(7, 0) == 0 and lib.mlb_fft_mixed_compiled(16384, 1) == 0
