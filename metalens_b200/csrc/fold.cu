// Aperture fold for FFT-bin-stride far-field grids (HBM-bound streaming kernel).
//
//   F[s*q] = sum_{m'} J[m'] e^{-2 pi i q (m'+h)/K}  (K = M/s)
//          = sum_{p<K} e^{-2 pi i q p / K} * G[p],   G[p] = sum_t J[((p-h) mod K) + t*K]
// so the M-point sum of nearfield_farfield.py:111-116 restricted to every s-th bin is a
// K-point sum of the folded aperture G.  Each input sample is read exactly once.
#include "common.cuh"

namespace mlb {

struct FoldArgs {
    const float2 *J[4];
    float2 *G[4];
    int ldj, ldg, K1, K2, s1, s2, h1, h2;
};

template <int VEC>
__global__ void __launch_bounds__(128) fold_kernel(FoldArgs a) {
    const int p2 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int p1 = blockIdx.y;
    if (p2 >= a.K2) return;
    const float2 *__restrict__ J = pick4(a.J, blockIdx.z);
    int m1 = (p1 - a.h1) % a.K1;
    if (m1 < 0) m1 += a.K1;
    int m2 = (p2 - a.h2) % a.K2;
    if (m2 < 0) m2 += a.K2;
    float acc[2 * VEC];
#pragma unroll
    for (int v = 0; v < 2 * VEC; ++v) acc[v] = 0.f;
    for (int t1 = 0; t1 < a.s1; ++t1) {
        const float2 *row = J + (size_t)(m1 + t1 * a.K1) * a.ldj + m2;
#pragma unroll 4
        for (int t2 = 0; t2 < a.s2; ++t2) {
            if (VEC == 2) {
                float4 v = __ldcs(reinterpret_cast<const float4 *>(row + (size_t)t2 * a.K2));
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            } else {
                float2 v = __ldcs(row + (size_t)t2 * a.K2);
                acc[0] += v.x; acc[1] += v.y;
            }
        }
    }
    float2 *g = pick4(a.G, blockIdx.z) + (size_t)p1 * a.ldg + p2;
    if (VEC == 2)
        *reinterpret_cast<float4 *>(g) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else
        *g = make_float2(acc[0], acc[1]);
}

}  // namespace mlb

extern "C" int mlb_fold(const mlb_c64 *const *h_J, int ldj, int M1, int M2, int s1, int s2, int h1, int h2,
                        mlb_c64 *const *h_G, int ldg, int batch, void *stream) {
    MLB_REQUIRE(h_J && h_G && batch >= 1 && batch <= 4, "mlb_fold: bad batch %d", batch);
    MLB_REQUIRE(s1 >= 1 && s2 >= 1 && M1 % s1 == 0 && M2 % s2 == 0, "mlb_fold: stride (%d,%d) must divide (%d,%d)",
                s1, s2, M1, M2);
    mlb::FoldArgs a;
    a.K1 = M1 / s1; a.K2 = M2 / s2; a.s1 = s1; a.s2 = s2; a.h1 = h1; a.h2 = h2; a.ldj = ldj; a.ldg = ldg;
    MLB_REQUIRE(ldj >= M2 && ldg >= a.K2, "mlb_fold: leading dimensions too small");
    bool vec = (a.K2 % 2 == 0) && (h2 % 2 == 0) && (ldj % 2 == 0) && (ldg % 2 == 0);
    for (int b = 0; b < 4; ++b) {
        a.J[b] = reinterpret_cast<const float2 *>(h_J[b < batch ? b : 0]);
        a.G[b] = reinterpret_cast<float2 *>(h_G[b < batch ? b : 0]);
        vec = vec && mlb::aligned16(a.J[b]) && mlb::aligned16(a.G[b]);
    }
    dim3 block(128);
    if (vec) {
        dim3 grid((a.K2 / 2 + 127) / 128, a.K1, batch);
        mlb::fold_kernel<2><<<grid, block, 0, (cudaStream_t)stream>>>(a);
    } else {
        dim3 grid((a.K2 + 127) / 128, a.K1, batch);
        mlb::fold_kernel<1><<<grid, block, 0, (cudaStream_t)stream>>>(a);
    }
    return mlb::check_launch("mlb_fold");
}
