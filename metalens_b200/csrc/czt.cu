// Chirp-z (Bluestein) aperture sum for UNIFORM direction-cosine grids that are not FFT bins ("zoomed" far fields).
//
// The reference's grid is forced to the FFT bins (nearfield_farfield.py:35-39); its derivation (:97-120) holds for any
// (ux, uy).  For ux_i = u0 + i du and aperture coordinates x_m = (m - o) d the kernel factorises,
//     e^{-ik x_m ux_i} = e^{-i pi (t1 m' + tq m'^2 / 2)} . e^{+i pi tq (i - m')^2 / 2} . e^{-i pi tq i^2 / 2},
//     m' = m - o,  t1 = 2 n d u0 / lambda,  tq = 2 n d du / lambda,
// so the sum over m is a linear convolution with a chirp, done with the FFT passes of fft.cu at length
// L >= M + K - 1 (a power of two <= 8192): ~2 L log L work per row instead of M K, and float32-exact (every chirp phase
// is formed in float64 and reduced exactly by sincospi before the single rounding).  This file holds the chirp tables
// and the three pointwise kernels between the FFT passes; metalens_b200/farfield.py (method 'czt') sequences them.
#include "common.cuh"

namespace mlb {

__device__ __forceinline__ float2 cmul_f(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// pre[m] = e^{-i pi (t1 m' + tq m'^2/2)} (m < M);  kern[(n mod L)] = e^{+i pi tq n^2/2} for n in [o-M+1, K-1+o];
// post[i] = e^{-i pi tq i^2/2} / L (i < K)
__global__ void czt_chirps_kernel(int M, int o, int K, int L, double t1, double tq, float2 *__restrict__ pre,
                                  float2 *__restrict__ kern, float2 *__restrict__ post) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double s, c;
    if (t < M) {
        const double mp = (double)(t - o);
        sincospi(-(t1 * mp + 0.5 * tq * mp * mp), &s, &c);
        pre[t] = make_float2((float)c, (float)s);
    }
    if (t < L) {
        // slot t holds n = t for t <= K-1+o, n = t - L for the wrapped negative lags, 0 elsewhere
        float2 v = make_float2(0.f, 0.f);
        int n = t;
        bool used = (t <= K - 1 + o);
        if (!used && t - L >= o - M + 1) { n = t - L; used = true; }
        if (used) {
            const double nn = (double)n;
            sincospi(0.5 * tq * nn * nn, &s, &c);
            v = make_float2((float)c, (float)s);
        }
        kern[t] = v;
    }
    if (t < K) {
        const double ii = (double)t;
        sincospi(-0.5 * tq * ii * ii, &s, &c);
        const double inv = 1.0 / (double)L;
        post[t] = make_float2((float)(c * inv), (float)(s * inv));
    }
}

struct CztArgs {
    const float2 *in[4];
    float2 *out[4];
    const float2 *row_tab, *col_tab;     // optional multipliers indexed by the output row / column (NULL = 1)
    int ld_in, ld_out, rows_out, cols_out, rows_valid, cols_valid, row_off, col_off, conj_in, conj_out;
};

// out[r][c] = conj?( conj?(in[r + row_off][c + col_off]) . row_tab[r] . col_tab[c] ) for r < rows_valid, c < cols_valid, else 0
__global__ void __launch_bounds__(256) czt_pointwise_kernel(const CztArgs a) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= a.cols_out) return;
    float2 v = make_float2(0.f, 0.f);
    if (r < a.rows_valid && c < a.cols_valid) {
        v = pick4(a.in, blockIdx.z)[(size_t)(r + a.row_off) * a.ld_in + c + a.col_off];
        if (a.conj_in) v.y = -v.y;
        if (a.row_tab) v = cmul_f(v, __ldg(a.row_tab + r));
        if (a.col_tab) v = cmul_f(v, __ldg(a.col_tab + c));
        if (a.conj_out) v.y = -v.y;
    }
    pick4(a.out, blockIdx.z)[(size_t)r * a.ld_out + c] = v;
}

}  // namespace mlb

extern "C" int mlb_czt_chirps(int M, int origin, int K, int L, double t_lin, double t_quad, mlb_c64 *pre, mlb_c64 *kern,
                              mlb_c64 *post, void *stream) {
    MLB_REQUIRE(pre && kern && post, "mlb_czt_chirps: NULL pointer");
    MLB_REQUIRE(M >= 1 && K >= 1 && L >= M + K - 1 && origin >= 0 && origin <= M, "mlb_czt_chirps: need L >= M + K - 1 (%d, %d, %d)",
                M, K, L);
    const int n = L > M ? (L > K ? L : K) : (M > K ? M : K);
    mlb::czt_chirps_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(M, origin, K, L, t_lin, t_quad,
                                                                              reinterpret_cast<float2 *>(pre),
                                                                              reinterpret_cast<float2 *>(kern),
                                                                              reinterpret_cast<float2 *>(post));
    return mlb::check_launch("mlb_czt_chirps");
}

extern "C" int mlb_czt_pointwise(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int rows_out,
                                 int cols_out, int rows_valid, int cols_valid, int row_off, int col_off,
                                 const mlb_c64 *row_tab, const mlb_c64 *col_tab, int conj_in, int conj_out, int batch,
                                 void *stream) {
    MLB_REQUIRE(h_in && h_out && batch >= 1 && batch <= 4, "mlb_czt_pointwise: bad batch %d", batch);
    MLB_REQUIRE(rows_out > 0 && cols_out > 0 && rows_valid >= 0 && cols_valid >= 0 && rows_valid <= rows_out &&
                    cols_valid <= cols_out && ld_out >= cols_out && ld_in >= cols_valid + col_off && row_off >= 0 && col_off >= 0,
                "mlb_czt_pointwise: bad sizes");
    MLB_REQUIRE(rows_out <= 65535, "mlb_czt_pointwise: too many rows");
    mlb::CztArgs a;
    for (int b = 0; b < 4; ++b) {
        const int s = b < batch ? b : 0;
        MLB_REQUIRE(h_in[s] && h_out[s], "mlb_czt_pointwise: NULL operand %d", s);
        a.in[b] = reinterpret_cast<const float2 *>(h_in[s]);
        a.out[b] = reinterpret_cast<float2 *>(h_out[s]);
    }
    a.row_tab = reinterpret_cast<const float2 *>(row_tab);
    a.col_tab = reinterpret_cast<const float2 *>(col_tab);
    a.ld_in = ld_in; a.ld_out = ld_out; a.rows_out = rows_out; a.cols_out = cols_out; a.rows_valid = rows_valid;
    a.cols_valid = cols_valid; a.row_off = row_off; a.col_off = col_off; a.conj_in = conj_in ? 1 : 0; a.conj_out = conj_out ? 1 : 0;
    dim3 grid((cols_out + 255) / 256, rows_out, batch);
    mlb::czt_pointwise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return mlb::check_launch("mlb_czt_pointwise");
}
