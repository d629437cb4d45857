// Twiddle tables exp(i*pi*scale*coord[m]*u[i]) with float64 phases.
#include "common.cuh"

namespace mlb {

__global__ void twiddle_kernel(const double *__restrict__ coord, int n_coord, const double *__restrict__ u,
                               int n_u, double scale, float2 *__restrict__ out, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;  // fast index: direction cosine
    int m = blockIdx.y;
    if (i >= n_u || m >= n_coord) return;
    // phase/pi in float64; sincospi() reduces the argument exactly, so |phase| ~ 1e4 rad
    // (k x' ux at M = 8192) costs no accuracy before the single rounding to fp32.
    double t = scale * coord[m] * u[i];
    double s, c;
    sincospi(t, &s, &c);
    out[(size_t)m * ld + i] = make_float2((float)c, (float)s);
}

}  // namespace mlb

extern "C" int mlb_twiddle_build(const double *coord, int n_coord, const double *u, int n_u, double scale,
                                 mlb_c64 *out, int ld, void *stream) {
    MLB_REQUIRE(coord && u && out, "mlb_twiddle_build: NULL pointer");
    MLB_REQUIRE(n_coord > 0 && n_u > 0 && ld >= n_u, "mlb_twiddle_build: bad sizes (%d,%d,ld=%d)", n_coord, n_u,
                ld);
    dim3 block(128), grid((n_u + 127) / 128, n_coord);
    mlb::twiddle_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(coord, n_coord, u, n_u, scale,
                                                                  reinterpret_cast<float2 *>(out), ld);
    return mlb::check_launch("mlb_twiddle_build");
}
