"""CPU: index bookkeeping of the one-aperture-over-G-ranks far field (metalens_b200/slab.py) emulated in numpy:
each rank folds and row-transforms only ITS aperture rows, scatters column slabs, every rank column-transforms its
slab -- the assembled result must equal fftshift(fft2(fftshift(J)))[::s, ::s] (nearfield_farfield.py:18-20, :68)."""
import numpy as np
import pytest

from metalens_b200.slab import slab_geometry, slab_rows


@pytest.mark.parametrize("M,s,world", [(64, 4, 1), (64, 4, 2), (64, 4, 4), (128, 2, 8), (96, 1, 2), (64, 4, 8)])
def test_distributed_fold_fft_equals_reference_bins(M, s, world):
    rng = np.random.default_rng(M + s + world)
    J = rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M))
    ref = np.fft.fftshift(np.fft.fft2(np.fft.fftshift(J)))[::s, ::s]
    K = M // s
    W = [np.zeros((K, K // world), complex) for _ in range(world)]            # rank p: all rows of its column slab
    owned = np.zeros(M, int)
    for rank in range(world):
        g = slab_geometry(M, M, s, s, rank, world)
        rows = g["x_rows"]
        owned[rows] += 1
        local = J[rows]                                                       # what the rank assembles / holds
        n = g["rows_per_rank"]
        assert np.array_equal(rows[:n] % K, rows[:n]) and np.array_equal(rows[n:2 * n], rows[:n] + K) or s == 1
        p = np.arange(K)
        for r in range(n):
            folded = np.zeros(K, complex)
            for t1 in range(s):
                for t2 in range(s):
                    folded += local[r + t1 * n][((p - g["roll_c"]) % K) + t2 * K]
            row = np.roll(np.fft.fft(folded), g["out_roll_rows"])
            R = (g["out_row0"] + (r // g["row_block"]) * g["row_stride"] + r % g["row_block"]) % K
            for peer in range(world):
                c0 = peer * g["cols_per_rank"]
                W[peer][R] = row[c0:c0 + g["cols_per_rank"]]
    assert (owned == 1).all()                                                  # every aperture row assembled exactly once
    F = np.concatenate([np.roll(np.fft.fft(w, axis=0), g["out_roll_cols"], axis=0) for w in W], axis=1)
    assert np.allclose(F, ref, rtol=1e-12, atol=1e-10 * np.abs(ref).max())


def test_slab_rows_layout():
    rows = slab_rows(64, 4, 1, 2)          # K1 = 16, 8 rows per rank in blocks of 4 dealt round-robin: rank 1 -> 4..7, 12..15
    assert rows.tolist() == [4, 5, 6, 7, 12, 13, 14, 15, 20, 21, 22, 23, 28, 29, 30, 31,
                             36, 37, 38, 39, 44, 45, 46, 47, 52, 53, 54, 55, 60, 61, 62, 63]
    assert slab_rows(32, 4, 1, 2).tolist() == [4, 5, 6, 7, 12, 13, 14, 15, 20, 21, 22, 23, 28, 29, 30, 31]
