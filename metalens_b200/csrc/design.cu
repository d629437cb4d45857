// Lens layout on the device: the hex-lattice centre of design_collimator.design_center
// (design_collimator.py:74-137) and the bin grid the assembly kernel searches for the nearest cell.
//
// The reference builds the lattice with Python loops over ~10^6 cells and picks every cell's HexGridSet entry with
// HexGridSet.pick_from_phase (lens_center.py:175-186).  Here: one thread per lattice column n2 counts its cells inside
// the circle (same float64 test x^2 + y^2 < radius^2), an exclusive scan of the counts gives every column its rows, and
// a second pass writes [x, y, index] in the reference's row order (n2 outer, n1 inner) -- no host sort, no upload.
#include "common.cuh"

namespace mlb {

struct HexArgs {
    double pitch, radius2, kwave, source_distance;
    int n1_lo, n1_hi, n2_lo, n2_hi;            // inclusive candidate ranges of design_collimator.hexagonal_grid
};

// lattice point of (n2, n1), the expressions of design_collimator.py:104-107
__device__ __forceinline__ void hex_point(const HexArgs &a, int n2, int n1, double &x, double &y) {
    x = a.pitch * (double)n2 * 1.7320508075688772 / 2.0;       // n * n2 * 3**0.5 / 2
    y = a.pitch * ((double)n1 + (double)n2 / 2.0);             // n * (n1 + n2 / 2)
}

__global__ void hex_count_kernel(const HexArgs a, int *__restrict__ counts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > a.n2_hi - a.n2_lo) return;
    const int n2 = a.n2_lo + c;
    int n = 0;
    for (int n1 = a.n1_lo; n1 <= a.n1_hi; ++n1) {
        double x, y;
        hex_point(a, n2, n1, x, y);
        n += (__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)) < a.radius2) ? 1 : 0;
    }
    counts[c] = n;
}

// offsets[c] = number of cells in the columns before c (exclusive scan of counts)
__global__ void hex_fill_kernel(const HexArgs a, const long long *__restrict__ offsets, const double2 *__restrict__ amp,
                                int n_amp, double *__restrict__ cells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > a.n2_hi - a.n2_lo) return;
    const int n2 = a.n2_lo + c;
    long long row = offsets[c];
    const double TWO_PI = 6.283185307179586, PI = 3.141592653589793;
    for (int n1 = a.n1_lo; n1 <= a.n1_hi; ++n1) {
        double x, y;
        hex_point(a, n2, n1, x, y);
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        if (!(r2 < a.radius2)) continue;
        // target_phase(r) + pi, design_collimator.py:57-60, :130: (-k (sqrt(f^2 + r^2) - f)) % 2 pi + pi
        const double r = sqrt(r2);
        const double raw = -a.kwave * (sqrt(__dadd_rn(__dmul_rn(a.source_distance, a.source_distance), __dmul_rn(r, r))) -
                                       a.source_distance);
        double m = fmod(raw, TWO_PI);
        if (m < 0.0) m += TWO_PI;                              // Python's % has the sign of the divisor
        const double phase = m + PI;
        double s, co;
        sincos(phase, &s, &co);
        // argmax_k Im(x_amp[k] e^{-i phase}) = Im(x_amp) cos - Re(x_amp) sin; first maximum wins (numpy argmax)
        int best = 0;
        double best_v = -1e300;
        for (int k = 0; k < n_amp; ++k) {
            const double2 v = amp[k];
            const double fom = v.y * co - v.x * s;
            if (fom > best_v) { best_v = fom; best = k; }
        }
        cells[3 * row] = x; cells[3 * row + 1] = y; cells[3 * row + 2] = (double)best;
        ++row;
    }
}

struct BinArgs {
    const double *cells;                     // [n][3]
    int n, nbx, nby;
    double x0, y0, inv_size;
};
__device__ __forceinline__ int bin_key(const BinArgs &a, double x, double y) {
    int bx = (int)floor((x - a.x0) * a.inv_size), by = (int)floor((y - a.y0) * a.inv_size);
    bx = min(max(bx, 0), a.nbx - 1);
    by = min(max(by, 0), a.nby - 1);
    return by * a.nbx + bx;
}
__global__ void cells_bin_count_kernel(const BinArgs a, int *__restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) atomicAdd(count + bin_key(a, a.cells[3 * i], a.cells[3 * i + 1]), 1);
}
// cursor[b] starts at bin_start[b]; the order inside a bin is arbitrary (the assembly kernel breaks exact distance ties
// by the ORIGINAL row, so its result does not depend on it)
__global__ void cells_bin_scatter_kernel(const BinArgs a, int *__restrict__ cursor, double *__restrict__ cell_x,
                                         double *__restrict__ cell_y, int *__restrict__ cell_which,
                                         int *__restrict__ cell_orig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double x = a.cells[3 * i], y = a.cells[3 * i + 1];
    const int pos = atomicAdd(cursor + bin_key(a, x, y), 1);
    cell_x[pos] = x; cell_y[pos] = y;
    cell_which[pos] = (int)(long long)a.cells[3 * i + 2];         // .astype(int), nearfield.py:367
    cell_orig[pos] = i;
}

}  // namespace mlb

static int fill_hex(mlb::HexArgs &a, double pitch, double radius, double wavelength, double refractive_index,
                    double source_distance, int n1_lo, int n1_hi, int n2_lo, int n2_hi) {
    MLB_REQUIRE(pitch > 0 && radius > 0 && n1_hi >= n1_lo && n2_hi >= n2_lo, "mlb_hex_*: bad lattice");
    a.pitch = pitch; a.radius2 = radius * radius;
    a.kwave = 2.0 * 3.141592653589793 * refractive_index / wavelength;      // design_collimator.py:58
    a.source_distance = source_distance;
    a.n1_lo = n1_lo; a.n1_hi = n1_hi; a.n2_lo = n2_lo; a.n2_hi = n2_hi;
    return MLB_OK;
}

extern "C" int mlb_hex_count(double pitch, double radius, int n1_lo, int n1_hi, int n2_lo, int n2_hi, int *counts,
                             void *stream) {
    mlb::HexArgs a;
    if (int rc = fill_hex(a, pitch, radius, 1.0, 1.0, 1.0, n1_lo, n1_hi, n2_lo, n2_hi)) return rc;
    MLB_REQUIRE(counts, "mlb_hex_count: NULL counts");
    const int cols = n2_hi - n2_lo + 1;
    mlb::hex_count_kernel<<<(cols + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a, counts);
    return mlb::check_launch("mlb_hex_count");
}

extern "C" int mlb_hex_fill(double pitch, double radius, int n1_lo, int n1_hi, int n2_lo, int n2_hi,
                            const long long *offsets, double wavelength, double refractive_index,
                            double source_distance, const double *x_amp, int n_amp, double *cells, void *stream) {
    mlb::HexArgs a;
    if (int rc = fill_hex(a, pitch, radius, wavelength, refractive_index, source_distance, n1_lo, n1_hi, n2_lo, n2_hi)) return rc;
    MLB_REQUIRE(offsets && x_amp && cells && n_amp >= 1 && wavelength > 0, "mlb_hex_fill: bad arguments");
    const int cols = n2_hi - n2_lo + 1;
    mlb::hex_fill_kernel<<<(cols + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a, offsets, reinterpret_cast<const double2 *>(x_amp),
                                                                           n_amp, cells);
    return mlb::check_launch("mlb_hex_fill");
}

extern "C" int mlb_cells_bin(const double *cells, int n, double x0, double y0, double bin_size, int nbx, int nby,
                             int phase, int *count_or_cursor, double *cell_x, double *cell_y, int *cell_which,
                             int *cell_orig, void *stream) {
    MLB_REQUIRE(cells && n > 0 && bin_size > 0 && nbx > 0 && nby > 0 && count_or_cursor, "mlb_cells_bin: bad arguments");
    mlb::BinArgs a;
    a.cells = cells; a.n = n; a.nbx = nbx; a.nby = nby; a.x0 = x0; a.y0 = y0; a.inv_size = 1.0 / bin_size;
    const int blocks = (n + 255) / 256;
    if (phase == 0) {
        mlb::cells_bin_count_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, count_or_cursor);
    } else {
        MLB_REQUIRE(cell_x && cell_y && cell_which && cell_orig, "mlb_cells_bin: NULL outputs");
        mlb::cells_bin_scatter_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, count_or_cursor, cell_x, cell_y, cell_which, cell_orig);
    }
    return mlb::check_launch("mlb_cells_bin");
}
