// Radiated-power epilogue: aperture sums (Ex,Ey,Hx,Hy) -> P(ux,uy).
// Follows farfield_from_nearfield_helper (reference nearfield_farfield.py:135-189) in
// float64 arithmetic; K^2 points x ~100 flop is negligible next to the aperture sum.
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"

namespace mlb {

struct EpiArgs {
    const float2 *F[4];   // Ex, Ey, Hx, Hy aperture sums
    const double *ux, *uy;
    void *P;
    double *block_sums;
    double amp_scale, pref, Z;
    int ldf, ldp, Kx, Ky, p_is_double, accumulate;
};

struct cd { double re, im; };
__device__ __forceinline__ cd cmul(cd a, double s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cd csub(cd a, cd b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ double cabs2(cd a) { return a.re * a.re + a.im * a.im; }

constexpr int EPI_THREADS = 256;

struct cfl { float re, im; };
__device__ __forceinline__ cfl fmul(cfl a, float s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ cfl fadd(cfl a, cfl b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cfl fsub(cfl a, cfl b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ float fabs2(cfl a) { return a.re * a.re + a.im * a.im; }

// F32: the float32-output flavour.  The evanescent mask and the DC test stay in float64 (bit-identical
// to the reference); the polar projections and |.|^2 run in fp32 (relative error ~1e-7, the inputs are
// complex64 anyway), which makes the kernel memory- instead of FP64-pipe-bound.
template <bool F32, bool IN64 = false>
__global__ void __launch_bounds__(EPI_THREADS) ff_epilogue_kernel(EpiArgs a) {
    const long long n = (long long)blockIdx.x * EPI_THREADS + threadIdx.x;
    const long long total = (long long)a.Kx * a.Ky;
    double p = 0.0;
    bool finite = false;
    if (n < total) {
        const int i = (int)(n / a.Ky), j = (int)(n % a.Ky);
        const double ux = a.ux[i], uy = a.uy[j];
        const size_t off = (size_t)i * a.ldf + j;
        // IN64: the caller's complex128 aperture sums as they are (the strict drop-in, nearfield_farfield.py:14)
        using in_t = typename std::conditional<IN64, double2, float2>::type;
        const in_t fex = reinterpret_cast<const in_t *>(a.F[0])[off], fey = reinterpret_cast<const in_t *>(a.F[1])[off];
        const in_t fhx = reinterpret_cast<const in_t *>(a.F[2])[off], fhy = reinterpret_cast<const in_t *>(a.F[3])[off];
        const double s_ = a.amp_scale;
        // (8.15) J = n x H, M = -n x E with n = +z  (reference :135-138)
        const cd Nx = {-(double)fhy.x * s_, -(double)fhy.y * s_};
        const cd Ny = {(double)fhx.x * s_, (double)fhx.y * s_};
        const cd Lx = {(double)fey.x * s_, (double)fey.y * s_};
        const cd Ly = {-(double)fex.x * s_, -(double)fex.y * s_};
        // uz^2 exactly as numpy evaluates (1 - ux**2 - uy**2): no FMA contraction, so the
        // evanescent (NaN) mask is bit-identical to the reference's (:153-155).
        const double ux2 = __dmul_rn(ux, ux), uy2 = __dmul_rn(uy, uy);
        const double uz2 = __dsub_rn(__dsub_rn(1.0, ux2), uy2);
        if (F32) {
            // raw aperture sums here; the common factor amp_scale^2 is applied in float64 at the end, so the
            // fp32 part never sees the (unit-system dependent) dx*dy scale
            const cfl nx = {-(float)fhy.x, -(float)fhy.y}, ny = {(float)fhx.x, (float)fhx.y};
            const cfl lx = {(float)fey.x, (float)fey.y}, ly = {-(float)fex.x, -(float)fex.y};
            cfl nth, nph, lth, lph;
            float uzf;
            if (uz2 < 0.0) {
                uzf = CUDART_NAN_F;
                nth = nph = lth = lph = {CUDART_NAN_F, CUDART_NAN_F};
            } else {
                uzf = sqrtf((float)uz2);
                if (ux == 0.0 && uy == 0.0) {
                    nth = nx; nph = ny; lth = lx; lph = ly;
                } else {
                    const float inv = 1.0f / ((float)sqrt(__dadd_rn(ux2, uy2)) + 1e-9f);
                    const float px = (float)ux * inv, py = (float)uy * inv, cx = px * uzf, cy = py * uzf;
                    nth = fadd(fmul(nx, cx), fmul(ny, cy));
                    nph = fsub(fmul(ny, px), fmul(nx, py));
                    lth = fadd(fmul(lx, cx), fmul(ly, cy));
                    lph = fsub(fmul(ly, px), fmul(lx, py));
                }
            }
            const float Zf = (float)a.Z;
            const float mag = fabs2(fadd(lph, fmul(nth, Zf))) + fabs2(fsub(lth, fmul(nph, Zf)));
            p = a.pref * s_ * s_ * (double)mag / ((double)uzf + 1e-5) * 2.0;
            float pf = (float)p;
            float *dst = reinterpret_cast<float *>(a.P) + (size_t)i * a.ldp + j;
            if (a.accumulate) pf += *dst;                   // incoherent sum over sources
            *dst = pf;
            finite = isfinite(pf);
        } else {
        const double uz = (uz2 < 0.0) ? CUDART_NAN : sqrt(uz2);
        const double sinth = sqrt(__dadd_rn(ux2, uy2));
        const double d = sinth + 1e-9;                                    // :158 regulariser
        cd Nth, Nph, Lth, Lph;
        if (ux == 0.0 && uy == 0.0) {                                     // :161-169
            Nth = Nx; Nph = Ny; Lth = Lx; Lph = Ly;
        } else {
            const double cx = ux * uz / d, cy = uy * uz / d, px = ux / d, py = uy / d;
            Nth = cadd(cmul(Nx, cx), cmul(Ny, cy));                       // :158
            Nph = csub(cmul(Ny, px), cmul(Nx, py));                       // :159
            Lth = cadd(cmul(Lx, cx), cmul(Ly, cy));                       // :166
            Lph = csub(cmul(Ly, px), cmul(Lx, py));                       // :167
        }
        const cd t1 = cadd(Lph, cmul(Nth, a.Z));
        const cd t2 = csub(Lth, cmul(Nph, a.Z));
        p = a.pref * (cabs2(t1) + cabs2(t2)) / (uz + 1e-5) * 2.0;         // :184-189
        const size_t po = (size_t)i * a.ldp + j;
        if (a.p_is_double) {
            double *dst = reinterpret_cast<double *>(a.P) + po;
            *dst = a.accumulate ? *dst + p : p;
        } else {
            float *dst = reinterpret_cast<float *>(a.P) + po;
            *dst = a.accumulate ? *dst + (float)p : (float)p;
        }
        finite = isfinite(p);
        }
    }
    if (a.block_sums) {                                                   // :74 total_P over finite bins
        double v = finite ? p : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __shared__ double ws[EPI_THREADS / 32];
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < EPI_THREADS / 32; ++w) s += ws[w];
            a.block_sums[blockIdx.x] = s;
        }
    }
}

// Two adjacent far-field points per thread (16-byte loads of the four aperture sums, 8-byte store of P where the
// pitch of P allows): the float32-output flavour for even-pitched aperture sums, any row length.  Same arithmetic as ff_epilogue_kernel<true> except that
// sin(theta) and the final quotient are formed in fp32 (relative error ~1e-7); the evanescent mask, the DC test
// and the unit-system dependent scale stay in float64.  128 threads x 2 points per block: mlb_ff_epilogue_blocks()
// counts pairs per row, so every flavour launches (and fills block_sums for) the same number of blocks.
__global__ void __launch_bounds__(EPI_THREADS / 2) ff_epilogue_x2_kernel(EpiArgs a, int vec_store) {
    // pair n covers points (i, j) and (i, j + 1), j even; an odd row length leaves the last pair of a row half empty
    // (the loads still hit the even-pitched buffers, the second point is neither stored nor summed)
    const int ppr = (a.Ky + 1) >> 1;
    const long long n = (long long)blockIdx.x * (EPI_THREADS / 2) + threadIdx.x;
    const long long total = (long long)a.Kx * ppr;
    double sum = 0.0;
    if (n < total) {
        const int i = (int)(n / ppr), j = (int)(n % ppr) * 2;
        const bool second = j + 1 < a.Ky;
        const double ux = a.ux[i];
        const double2 uy2v = make_double2(a.uy[j], second ? a.uy[j + 1] : 0.0);
        const size_t off = (size_t)i * a.ldf + j;
        const float4 fex = *reinterpret_cast<const float4 *>(a.F[0] + off), fey = *reinterpret_cast<const float4 *>(a.F[1] + off);
        const float4 fhx = *reinterpret_cast<const float4 *>(a.F[2] + off), fhy = *reinterpret_cast<const float4 *>(a.F[3] + off);
        const double ux2 = __dmul_rn(ux, ux);
        const float Zf = (float)a.Z;
        const double scale = a.pref * a.amp_scale * a.amp_scale * 2.0;
        float *dst = reinterpret_cast<float *>(a.P) + (size_t)i * a.ldp + j;
        float pf[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double uy = q ? uy2v.y : uy2v.x;
            const cfl ex = q ? cfl{fex.z, fex.w} : cfl{fex.x, fex.y}, ey = q ? cfl{fey.z, fey.w} : cfl{fey.x, fey.y};
            const cfl hx = q ? cfl{fhx.z, fhx.w} : cfl{fhx.x, fhx.y}, hy = q ? cfl{fhy.z, fhy.w} : cfl{fhy.x, fhy.y};
            const double uyy = __dmul_rn(uy, uy);
            const double uz2 = __dsub_rn(__dsub_rn(1.0, ux2), uyy);              // as numpy: bit-identical NaN mask
            const float uzf = (uz2 < 0.0) ? CUDART_NAN_F : sqrtf((float)uz2);
            float px = 1.f, py = 0.f;                                            // DC bin: Cartesian components (:161-169)
            if (!(ux == 0.0 && uy == 0.0)) {
                const float inv = 1.0f / (sqrtf((float)__dadd_rn(ux2, uyy)) + 1e-9f);
                px = (float)ux * inv; py = (float)uy * inv;
            }
            const float cx = px * uzf, cy = py * uzf;
            // t1 = L_phi + Z N_theta, t2 = L_theta - Z N_phi with N = (-Fhy, Fhx), L = (Fey, -Fex)
            const cfl t1 = {-px * ex.re - py * ey.re + Zf * (cy * hx.re - cx * hy.re),
                            -px * ex.im - py * ey.im + Zf * (cy * hx.im - cx * hy.im)};
            const cfl t2 = {cx * ey.re - cy * ex.re - Zf * (px * hx.re + py * hy.re),
                            cx * ey.im - cy * ex.im - Zf * (px * hx.im + py * hy.im)};
            const float qf = (fabs2(t1) + fabs2(t2)) / (uzf + 1e-5f);
            const double p = scale * (double)qf;
            pf[q] = (float)p;
            if (q == 1 && !second) break;
            if (a.accumulate) pf[q] += dst[q];
            if (isfinite(pf[q])) sum += p;
        }
        if (second && vec_store) {
            *reinterpret_cast<float2 *>(dst) = make_float2(pf[0], pf[1]);
        } else {
            dst[0] = pf[0];
            if (second) dst[1] = pf[1];
        }
    }
    if (a.block_sums) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __shared__ double ws[EPI_THREADS / 64];
        if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < EPI_THREADS / 64; ++w) s += ws[w];
            a.block_sums[blockIdx.x] = s;
        }
    }
}

// Figure-of-merit reduction: per-block sums of P over (a) all finite bins and
// (b) the finite bins inside a cone (ux-ux0)^2 + (uy-uy0)^2 <= radius^2 around a target direction.
template <typename T>
__global__ void __launch_bounds__(EPI_THREADS) cone_power_kernel(const T *__restrict__ P, int ldp,
                                                                 const double *__restrict__ ux,
                                                                 const double *__restrict__ uy, int Kx, int Ky,
                                                                 double ux0, double uy0, double r2,
                                                                 double *__restrict__ cone_sums,
                                                                 double *__restrict__ total_sums) {
    const long long n = (long long)blockIdx.x * EPI_THREADS + threadIdx.x;
    double c = 0.0, t = 0.0;
    if (n < (long long)Kx * Ky) {
        const int i = (int)(n / Ky), j = (int)(n % Ky);
        const double p = (double)P[(size_t)i * ldp + j];
        if (isfinite(p)) {
            t = p;
            const double dx = ux[i] - ux0, dy = uy[j] - uy0;
            if (dx * dx + dy * dy <= r2) c = p;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __shared__ double wc[EPI_THREADS / 32], wt[EPI_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { wc[threadIdx.x >> 5] = c; wt[threadIdx.x >> 5] = t; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sc = 0.0, st = 0.0;
#pragma unroll
        for (int w = 0; w < EPI_THREADS / 32; ++w) { sc += wc[w]; st += wt[w]; }
        cone_sums[blockIdx.x] = sc;
        total_sums[blockIdx.x] = st;
    }
}

__global__ void __launch_bounds__(256) sum_f64_kernel(const double *__restrict__ in, int n, double scale,
                                                      double *__restrict__ out) {
    __shared__ double ws[8];
    const double s = ordered_sum_256(in, n, ws);
    if (threadIdx.x == 0) out[0] = s * scale;
}

}  // namespace mlb

extern "C" int mlb_ff_epilogue_blocks(int Kx, int Ky) {
    long long t = (long long)Kx * ((Ky + 1) / 2) * 2;          // whole pairs per row (odd rows end in a half pair)
    return (int)((t + mlb::EPI_THREADS - 1) / mlb::EPI_THREADS);
}

extern "C" int mlb_ff_epilogue(const mlb_c64 *const *h_Fhat, int ldf, const double *ux, const double *uy, int Kx,
                               int Ky, double amp_scale, double wavelength, double n_glass, double Z0, void *P,
                               int ldp, int p_is_double, double *block_sums, void *stream) {
    MLB_REQUIRE(h_Fhat && ux && uy && P, "mlb_ff_epilogue: NULL pointer");
    MLB_REQUIRE(Kx > 0 && Ky > 0 && ldf >= Ky && ldp >= Ky, "mlb_ff_epilogue: bad sizes");
    MLB_REQUIRE(wavelength > 0 && n_glass > 0 && Z0 > 0, "mlb_ff_epilogue: bad physical constants");
    mlb::EpiArgs a;
    for (int f = 0; f < 4; ++f) {
        MLB_REQUIRE(h_Fhat[f] != nullptr, "mlb_ff_epilogue: field %d is NULL", f);
        a.F[f] = reinterpret_cast<const float2 *>(h_Fhat[f]);
    }
    a.ux = ux; a.uy = uy; a.P = P; a.block_sums = block_sums;
    a.amp_scale = amp_scale;
    a.Z = Z0 / n_glass;                                                    // :183
    const double pi = 3.14159265358979323846;
    const double k = 2 * pi * n_glass / wavelength;
    a.pref = k * k / (32 * pi * pi * a.Z);                                  // :184
    a.ldf = ldf; a.ldp = ldp; a.Kx = Kx; a.Ky = Ky; a.p_is_double = p_is_double & 1; a.accumulate = (p_is_double >> 1) & 1;
    const bool in64 = (p_is_double & 4) != 0;
    MLB_REQUIRE(!in64 || (p_is_double & 1), "mlb_ff_epilogue: complex128 aperture sums need the float64 output");
    p_is_double &= 1;
    if (in64) {
        mlb::ff_epilogue_kernel<false, true><<<mlb_ff_epilogue_blocks(Kx, Ky), mlb::EPI_THREADS, 0, (cudaStream_t)stream>>>(a);
        return mlb::check_launch("mlb_ff_epilogue(complex128)");
    }
    bool x2 = !p_is_double && ldf % 2 == 0;
    for (int f = 0; f < 4; ++f) x2 = x2 && mlb::aligned16(a.F[f]);
    const int vec_store = (ldp % 2 == 0 && (reinterpret_cast<uintptr_t>(P) & 7u) == 0) ? 1 : 0;
    if (p_is_double) mlb::ff_epilogue_kernel<false><<<mlb_ff_epilogue_blocks(Kx, Ky), mlb::EPI_THREADS, 0, (cudaStream_t)stream>>>(a);
    else if (x2) mlb::ff_epilogue_x2_kernel<<<mlb_ff_epilogue_blocks(Kx, Ky), mlb::EPI_THREADS / 2, 0, (cudaStream_t)stream>>>(a, vec_store);
    else mlb::ff_epilogue_kernel<true><<<mlb_ff_epilogue_blocks(Kx, Ky), mlb::EPI_THREADS, 0, (cudaStream_t)stream>>>(a);
    return mlb::check_launch("mlb_ff_epilogue");
}

extern "C" int mlb_cone_power(const void *P, int ldp, int p_is_double, const double *ux, const double *uy, int Kx,
                              int Ky, double ux0, double uy0, double radius, double *cone_block_sums,
                              double *total_block_sums, void *stream) {
    MLB_REQUIRE(P && ux && uy && cone_block_sums && total_block_sums, "mlb_cone_power: NULL pointer");
    MLB_REQUIRE(Kx > 0 && Ky > 0 && ldp >= Ky && radius >= 0, "mlb_cone_power: bad sizes");
    const int nb = mlb_ff_epilogue_blocks(Kx, Ky);
    if (p_is_double)
        mlb::cone_power_kernel<double><<<nb, mlb::EPI_THREADS, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const double *>(P), ldp, ux, uy, Kx, Ky, ux0, uy0, radius * radius, cone_block_sums,
            total_block_sums);
    else
        mlb::cone_power_kernel<float><<<nb, mlb::EPI_THREADS, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float *>(P), ldp, ux, uy, Kx, Ky, ux0, uy0, radius * radius, cone_block_sums,
            total_block_sums);
    return mlb::check_launch("mlb_cone_power");
}

extern "C" int mlb_sum_f64(const double *in, int n, double scale, double *out, void *stream) {
    MLB_REQUIRE(in && out && n >= 0, "mlb_sum_f64: bad arguments");
    mlb::sum_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(in, n, scale, out);
    return mlb::check_launch("mlb_sum_f64");
}
