// Aperture-field assembly: fused build_nearfield (reference nearfield.py:66-480).
//
// One thread per aperture sample (i = x index, j = y index, j fastest so the complex output rows
// are written coalesced).  Everything the reference does with ~40 full-array numpy passes, four
// scipy RegularGridInterpolator calls per diffraction order and a cKDTree query happens here in
// registers, in float64 (geometry and phases need it; B200 has a full-rate FP64 pipe
// for this amount of work).  Tables are a few hundred KB per collection and stay L1/L2 resident;
// neighbouring threads hit the same interpolation cell, so the gathers are mostly broadcasts.
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"

namespace mlb {

struct cplx { double re, im; };
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx operator*(cplx a, double s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ cplx operator*(double s, cplx a) { return {a.re * s, a.im * s}; }

// interval i with g[i] <= x < g[i+1], clipped to [0, n-2]  (scipy _rgi find_indices)
__device__ __forceinline__ int find_interval(const double *__restrict__ g, int n, double x) {
    if (!(x >= g[0])) return 0;
    if (x >= g[n - 1]) return n - 2;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= g[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// order-preserving map double -> int64 so that atomicMin/atomicMax on integers order doubles
__device__ __forceinline__ long long enc_f64(double x) {
    long long b = __double_as_longlong(x);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}

struct Interp3 {
    int i0, i1, i2;
    double t0, t1, t2;
};

// Per-ring record written by nearfield_prepare_kernel (layout documented at mlb_lens_desc.ring_aux): everything
// a periphery sample needs about its ring in ONE 64-byte line, all values formed with the reference's own float64
// expressions (so reading them is bit-identical to recomputing them per sample).
struct __align__(16) RingAux {
    double rc, gp, apg, lat, qx, qy, t2;
    int gc, i2;
};
// float32 screens: order-box extent (units of kvac) and the grating-copy index guard
struct __align__(16) RingAuxF {
    float fqx, fqy, inv_fqx, inv_fqy, inv_apg, guard, pad0, pad1;
};
static_assert(sizeof(RingAux) == 64 && sizeof(RingAuxF) == 32, "ring records");

// launch-uniform scalars, computed once on the host (IEEE float64, same expressions as nearfield.py:213, :262)
struct NfUniform {
    double kvac, kg, inv_kvac, inv_kg_n, Hcoef, lut_scale;
    float ng2, cf;
};

// cell and weight on a uniformly spaced axis: (u - first) * inv_step, clipped like scipy's find_indices
__device__ __forceinline__ void locate_uniform(double u, double first, double inv_step, int n, int &i, double &t) {
    const double f = (u - first) * inv_step;
    int c = (int)f;                                   // f >= 0 inside the bounds; clipping handles the rest
    c = min(max(c, 0), n - 2);
    i = c;
    t = f - (double)c;
}
// first two axes of the table (third: precomputed per ring, or located by the caller)
template <bool FAST>
__device__ __forceinline__ void locate2(const mlb_table_pack &p, double u0, double u1, Interp3 &q) {
    if (FAST && p.uniform01) {
        locate_uniform(u0, p.u0_first, p.u0_inv_step, p.n_ux, q.i0, q.t0);
        locate_uniform(u1, p.u1_first, p.u1_inv_step, p.n_uy, q.i1, q.t1);
        return;
    }
    const double *a0 = p.axes, *a1 = p.axes + p.n_ux;
    q.i0 = find_interval(a0, p.n_ux, u0);
    q.i1 = find_interval(a1, p.n_uy, u1);
    q.t0 = (u0 - a0[q.i0]) / (a0[q.i0 + 1] - a0[q.i0]);
    q.t1 = (u1 - a1[q.i1]) / (a1[q.i1 + 1] - a1[q.i1]);
}
__device__ __forceinline__ void locate3(const mlb_table_pack &p, double u2, Interp3 &q) {
    const double *a2 = p.axes + p.n_ux + p.n_uy;
    q.i2 = find_interval(a2, p.n_g, u2);
    q.t2 = (u2 - a2[q.i2]) / (a2[q.i2 + 1] - a2[q.i2]);
}

// trilinear gather of the 4 slots (x/ampfy, x/ampfx, y/ampfy, y/ampfx) of one order
__device__ __forceinline__ void gather4(const mlb_table_pack &p, int order, const Interp3 &q, cplx (&amp)[4]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) amp[s] = {0.0, 0.0};
    const size_t per_order = (size_t)p.n_ux * p.n_uy * p.n_g;
    const double2 *__restrict__ base = reinterpret_cast<const double2 *>(p.values) + (size_t)order * per_order * 4;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double wa = a ? q.t0 : 1.0 - q.t0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const double wb = b ? q.t1 : 1.0 - q.t1;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double wc = c ? q.t2 : 1.0 - q.t2;
                const double w = wa * wb * wc;
                const double2 *v = base + (((size_t)(q.i0 + a) * p.n_uy + (q.i1 + b)) * p.n_g + (q.i2 + c)) * 4;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const double2 x = __ldg(v + s);
                    amp[s].re += x.x * w;
                    amp[s].im += x.y * w;
                }
            }
        }
    }
}

// e^{i x}: float64 sincos for the complex128 output; for the complex64 output the argument is reduced
// to [-1/2, 1/2] turns in float64 (exact to ~1e-16 turns) and the sine/cosine taken by the special-function unit
// (abs. error 4e-7, well inside the complex64 tolerance of the path).
struct cf { float re, im; };
__device__ __forceinline__ cf operator+(cf a, cf b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cf operator*(cf a, cf b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cf operator*(cf a, float s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ cf operator*(float s, cf a) { return {a.re * s, a.im * s}; }

// 1/sqrt(v) for positive, fp32-representable v: fp32 seed (2^-22) + two Newton steps in float64 (-> ~1 ulp);
// a third of the instructions of sqrt + division
__device__ __forceinline__ double rsqrt_fast(double v) {
    double y = (double)rsqrtf((float)v);
    const double h = 0.5 * v;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

__device__ __forceinline__ cplx expi(double x) {
    double s, c;
    sincos(x, &s, &c);
    return {c, s};
}
__device__ __forceinline__ cf expi_fast(double x) {
    const double t = x * 0.15915494309189535;                 // x / 2 pi
    const double fr = t - rint(t);                            // [-1/2, 1/2] turns, exact to ~1e-16 turns
    // the special-function unit on the reduced angle: |error| <= 2^-21.4 (4e-7) on [-pi, pi], a fifth of the
    // instructions of the polynomial above and still 8x below the complex64 tolerance of the path
    const float a = (float)fr * 6.28318530717958647692f;
    return {__cosf(a), __sinf(a)};
}

// The interpolation cell of one sample in a float32 2-D slice [order][iu][iv][slot] of a pack -- the third table
// coordinate is fixed per sample: the ring's grating period (slice pre-interpolated per ring by
// nearfield_ring_tables_kernel) or, in the centre, the integer cell kind (a node of the axis, nearfield.py:411).  Every
// diffraction order of the sample reads the SAME cell (the interpolation point does not depend on the order, :293), so
// the corner offsets and the 4 bilinear weights are formed once per sample, not once per order.
struct CellF {
    int base, sA, sB;         // float4 offsets: first corner, stride of the ux axis, stride of the uy axis
    float w[4];               // weight of corner (a, b) at index 2a + b
};
__device__ __forceinline__ CellF make_cell(const Interp3 &q, int sA, int sB) {
    CellF cell;
    cell.sA = sA; cell.sB = sB;
    cell.base = q.i0 * sA + q.i1 * sB;
    const float t0 = (float)q.t0, t1 = (float)q.t1;
    cell.w[0] = (1.f - t0) * (1.f - t1); cell.w[1] = (1.f - t0) * t1;
    cell.w[2] = t0 * (1.f - t1);         cell.w[3] = t0 * t1;
    return cell;
}

// One diffraction order in fp32 (complex64-output path): gather the 4 slots of the cell, form
//   E_a += Z0 [ S_fy kx ky + S_fx (ky^2+kz^2) ] / (k_g kz n) * phase      (nearfield.py:312-327 / :426-441)
//   E_b += Z0 [ S_fy (-kx^2-kz^2) - S_fx kx ky ] / (k_g kz n) * phase
//   H_a += S_fy * phase ;  H_b += S_fx * phase        with S_f* = Hw_x a_x,f* + Hw_y a_y,f*
// nx, ny = k / kvac (so the products stay well inside the fp32 range); cfac = Z0 kvac / (k_g n).
__device__ __forceinline__ void order_fast(const float4 *__restrict__ tbl, int per_order2, int o, const CellF &cell,
                                           float Hw_x, float Hw_y, float nx, float ny, float ng2, float cfac, cf phase,
                                           cf &Ea, cf &Eb, cf &Ha, cf &Hb) {
    const float4 *__restrict__ v = tbl + ((size_t)o * per_order2 + cell.base);
    cf amp[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) amp[s] = {0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const float4 *c0 = v + (a ? cell.sA : 0) + (b ? cell.sB : 0);
            const float w = cell.w[2 * a + b];
            const float4 x0 = __ldg(c0), x1 = __ldg(c0 + 1);                       // slots (0,1) and (2,3)
            amp[0].re = fmaf(x0.x, w, amp[0].re); amp[0].im = fmaf(x0.y, w, amp[0].im);
            amp[1].re = fmaf(x0.z, w, amp[1].re); amp[1].im = fmaf(x0.w, w, amp[1].im);
            amp[2].re = fmaf(x1.x, w, amp[2].re); amp[2].im = fmaf(x1.y, w, amp[2].im);
            amp[3].re = fmaf(x1.z, w, amp[3].re); amp[3].im = fmaf(x1.w, w, amp[3].im);
        }
    const float nz2 = ng2 - nx * nx - ny * ny;                // > 0: the order propagates in air, n_glass > 1
    const float f = cfac * rsqrtf(nz2);                       // Z0 / (k_g kz n) in units of 1/kvac
    // P1 = S_fy phase, P2 = S_fx phase; the E terms are real combinations of them:
    //   E_a += f [ P1 nx ny + P2 (ny^2 + nz^2) ],   E_b += f [ -P1 (nx^2 + nz^2) - P2 nx ny ],   H_a += P1,  H_b += P2
    const cf Sfy = {fmaf(Hw_x, amp[0].re, Hw_y * amp[2].re), fmaf(Hw_x, amp[0].im, Hw_y * amp[2].im)};
    const cf Sfx = {fmaf(Hw_x, amp[1].re, Hw_y * amp[3].re), fmaf(Hw_x, amp[1].im, Hw_y * amp[3].im)};
    const cf P1 = {fmaf(Sfy.re, phase.re, -Sfy.im * phase.im), fmaf(Sfy.re, phase.im, Sfy.im * phase.re)};
    const cf P2 = {fmaf(Sfx.re, phase.re, -Sfx.im * phase.im), fmaf(Sfx.re, phase.im, Sfx.im * phase.re)};
    const float cxy = f * (nx * ny), cyy = f * fmaf(ny, ny, nz2), cxx = -f * fmaf(nx, nx, nz2);
    Ea.re = fmaf(P1.re, cxy, fmaf(P2.re, cyy, Ea.re)); Ea.im = fmaf(P1.im, cxy, fmaf(P2.im, cyy, Ea.im));
    Eb.re = fmaf(P1.re, cxx, fmaf(P2.re, -cxy, Eb.re)); Eb.im = fmaf(P1.im, cxx, fmaf(P2.im, -cxy, Eb.im));
    Ha.re += P1.re; Ha.im += P1.im;
    Hb.re += P2.re; Hb.im += P2.im;
}

// one diffraction order's contribution in float64 (nearfield.py:306-327 / :420-441), both incident
// polarisations and both amplitudes at once (same formulas as order_fast)
__device__ __forceinline__ void add_order(const cplx (&amp)[4], double Hw_x, double Hw_y, double kx, double ky,
                                          double kz, double inv_kg_n, double Z0, cplx phase, cplx &Ea, cplx &Eb,
                                          cplx &Ha, cplx &Hb) {
    const cplx Sfy = Hw_x * amp[0] + Hw_y * amp[2];
    const cplx Sfx = Hw_x * amp[1] + Hw_y * amp[3];
    const double f = Z0 * inv_kg_n / kz;
    const cplx ea = (Sfy * (kx * ky) + Sfx * (ky * ky + kz * kz)) * f;
    const cplx eb = (Sfy * (-kx * kx - kz * kz) + Sfx * (-kx * ky)) * f;
    Ea = Ea + ea * phase;
    Eb = Eb + eb * phase;
    Ha = Ha + Sfy * phase;
    Hb = Hb + Sfx * phase;
}

__device__ __forceinline__ bool out_of_bounds(const mlb_table_pack &p, double u0, double u1, double u2, bool check2) {
    return (u0 < p.bounds[0]) | (u0 > p.bounds[1]) | (u1 < p.bounds[2]) | (u1 > p.bounds[3]) |
           (check2 & ((u2 < p.bounds[4]) | (u2 > p.bounds[5])));
}

// per-(pack, order) count / min / max for the reference's error messages (slow path, want_stats)
__device__ __forceinline__ void record_stats(const mlb_table_pack &p, int order, double u0, double u1, double u2,
                                             long long *stats) {
    long long *s = stats + (size_t)(p.stats_slot + order) * MLB_STATS_PER_ORDER;
    atomicAdd(reinterpret_cast<unsigned long long *>(s), 1ULL);
    atomicMin(s + 1, enc_f64(u0)); atomicMax(s + 2, enc_f64(u0));
    atomicMin(s + 3, enc_f64(u1)); atomicMax(s + 4, enc_f64(u1));
    atomicMin(s + 5, enc_f64(u2)); atomicMax(s + 6, enc_f64(u2));
}

constexpr int NF_THREADS = 128;

struct NfOut {
    void *F[4];
    double *power_warp_sums;
    long long *stats;
    int *violation;
    int ld, out_is_double, lg_wy;     // lg_wy: log2 of the warp tile's y extent (5 = 32 x 1 ... 2 = 4 x 8)
    // exact nearest-cell ties: reported to (tie_count, tie_list[tie_capacity]) when tie_count != NULL;
    // fix-up launches (FIX): thread t re-assembles sample fix_samples[t] with the forced cell fix_cells[t]
    int *tie_count, *tie_list;
    int tie_capacity, n_fix;
    const int *fix_samples, *fix_cells;
};

// The diffraction-order loop of one sample (periphery: primed grating frame, nearfield.py:263-327; centre:
// lab frame about the hex cell, :389-441).  u0, u1 = incident direction cosines in that frame, (X, Y) = sample
// position relative to the grating / cell centre, qx, qy = reciprocal-lattice steps, u2 = third table coordinate.
// Orders that can propagate in air satisfy |u0 + ox qx/kvac| <= 1 and |u1 + oy qy/kvac| <= 1: only that (small)
// index box is visited, through the pack's dense (ox,oy) -> order map; an fp32 screen with margin picks the box,
// the exact float64 test of :279 / :398 decides.  FAST: fp32 order terms into fp32 accumulators (E, H of the
// complex64 output), in units where the incident weights Hw carry no dipole scale.
template <bool STATS, bool FAST, typename Acc, typename W>
__device__ __forceinline__ void order_loop(const mlb_table_pack &p, const NfUniform &U, double Z0, double u0, double u1,
                                           double u2, bool check2, bool have_q3, Interp3 q, double X, double Y,
                                           double qx, double qy, float fqx, float fqy, float inv_fqx, float inv_fqy,
                                           W Hw_x, W Hw_y, const NfOut &out, Acc &Ea, Acc &Eb, Acc &Ha, Acc &Hb,
                                           const float4 *__restrict__ tbl, int sA, int sB, int per_order2) {
    const float fu0 = (float)u0, fu1 = (float)u1;
    const int R = p.order_radius, Wd = 2 * R + 1;
    const int ox_lo = max(-R, (int)ceilf((-1.001f - fu0) * inv_fqx)), ox_hi = min(R, (int)floorf((1.001f - fu0) * inv_fqx));
    const int oy_lo = max(-R, (int)ceilf((-1.001f - fu1) * inv_fqy)), oy_hi = min(R, (int)floorf((1.001f - fu1) * inv_fqy));
    const double k0 = U.kvac * u0, k1 = U.kvac * u1, kv2 = U.kvac * U.kvac;
    bool located = false;
    CellF cell;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const float nxf = fmaf((float)ox, fqx, fu0);
        const int *__restrict__ map_row = p.order_map + (ox + R) * Wd + R;
        for (int oy = oy_lo; oy <= oy_hi; ++oy) {
            // fp32 screen of :279 / :398 (|error| < 2e-6): clearly evanescent orders never reach the float64 test
            const float nyf = fmaf((float)oy, fqy, fu1);
            const float s2 = fmaf(nxf, nxf, nyf * nyf);
            if (s2 > 1.00002f) continue;
            const int o = map_row[oy];
            if (o < 0) continue;                                                       // order not in the tables
            const double kx = k0 + ox * qx;                                            // :268 / :395
            const double ky = k1 + oy * qy;                                            // :269 / :396
            if (s2 < 0.99998f || kx * kx + ky * ky <= kv2) {                           // :279 / :398
                if (STATS) record_stats(p, o, u0, u1, u2, out.stats);
                if (!located) {
                    located = true;
                    if (out_of_bounds(p, u0, u1, u2, check2)) atomicOr(out.violation, 1);   // :294-305 / :412-419
                    locate2<FAST>(p, u0, u1, q);
                    if constexpr (FAST) cell = make_cell(q, sA, sB);
                    else if (!have_q3) locate3(p, u2, q);
                }
                // kz (:287), phase about the grating / cell centre (:291, :408-409), table gathers, accumulation
                const double phase_arg = kx * X + ky * Y;
                if constexpr (FAST) {
                    order_fast(tbl, per_order2, o, cell, Hw_x, Hw_y, (float)(kx * U.inv_kvac), (float)(ky * U.inv_kvac),
                               U.ng2, U.cf, expi_fast(phase_arg), Ea, Eb, Ha, Hb);
                } else {
                    const double kz = sqrt(U.kg * U.kg - kx * kx - ky * ky);
                    cplx amp[4];
                    gather4(p, o, q, amp);
                    add_order(amp, Hw_x, Hw_y, kx, ky, kz, U.inv_kg_n, Z0, expi(phase_arg), Ea, Eb, Ha, Hb);
                }
            }
        }
    }
}

__device__ __forceinline__ void store_c(void *base, size_t off, cplx v, int is_double) {
    if (is_double) reinterpret_cast<double2 *>(base)[off] = make_double2(v.re, v.im);
    else reinterpret_cast<float2 *>(base)[off] = make_float2((float)v.re, (float)v.im);
}

template <bool STATS, bool FAST, int MINB, bool FIX = false>
__global__ void __launch_bounds__(NF_THREADS, MINB) nearfield_kernel(const __grid_constant__ mlb_lens_desc L,
                                                                      const __grid_constant__ NfUniform U, const NfOut out) {
    using Acc = typename std::conditional<FAST, cf, cplx>::type;      // per-sample field accumulators
    using W = typename std::conditional<FAST, float, double>::type;   // incident weights
    // a warp covers wy (y, fast) x 32/wy (x) samples, the four warps of a block are stacked along y: compact warp
    // footprints touch fewer rings and table cells per gather than a 32 x 1 line that crosses the rings radially
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wy = 1 << out.lg_wy;
    int j = (blockIdx.x * (NF_THREADS / 32) + warp) * wy + (lane & (wy - 1));         // y index (fast)
    int i = blockIdx.y * (32 >> out.lg_wy) + (lane >> out.lg_wy);                     // x index
    int forced_cell = -1;
    if (FIX) {                                                                        // one listed sample per thread
        const int t = blockIdx.x * NF_THREADS + threadIdx.x;
        i = L.nx; j = L.ny;
        if (t < out.n_fix) {
            const int lin = out.fix_samples[t];
            i = lin / L.ny; j = lin - i * L.ny;
            forced_cell = out.fix_cells[t];
        }
    }
    const double PI = 3.14159265358979323846;
    double local_power = 0.0;
    if (j < L.ny && i < L.nx) {
        const double x = L.x_pts[i], y = L.y_pts[j];
        const double r = sqrt(x * x + y * y);                                  // nearfield.py:118
        // which_ring = searchsorted(boundaries, r) - 1  (left-biased, :125-128): the bin table brackets the
        // answer to the boundaries of three bins, the same bisection as before finishes it
        int b = (int)(r * U.lut_scale);
        b = min(max(b, 0), L.n_lut - 1);
        int lo = L.ring_lut[max(b - 1, 0)], hi = L.ring_lut[min(b + 2, L.n_lut)];
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (L.ring_boundary[mid] < r) lo = mid + 1; else hi = mid;
        }
        int ring = lo - 1;
        const bool in_center = (ring == -1);
        if (ring == L.n_rings) ring = -1;

        // incident direction and dipole field (:172-228).  FAST: one float64 reciprocal instead of four
        // divisions (<= 1 ulp each); the incident weights then leave out the dipole scale Hcoef, which multiplies
        // the float64 power term and the final fields instead, so fp32 never sees the unit system.
        double ux = 0.0, uy = 0.0, uz = 1.0, dHx, dHy, dEx, dEy, scale;
        if (L.plane_wave) {
            const double px = L.source_pol == 0 ? 1.0 : 0.0, py = L.source_pol == 1 ? 1.0 : 0.0;
            scale = FAST ? L.dipole_moment : 1.0;
            const double dm = FAST ? 1.0 : L.dipole_moment;
            dEx = px * dm; dEy = py * dm;
            dHx = -py * dm / L.Z0; dHy = px * dm / L.Z0;
        } else {
            const double dx = x - L.source_x, dy = y - L.source_y, dz = 0.0 - L.source_z;
            double a;
            if (FAST) {
                const double inv = rsqrt_fast(dx * dx + dy * dy + dz * dz);
                ux = dx * inv; uy = dy * inv; uz = dz * inv;
                a = uz * rsqrt_fast(uz) * inv;                                      // sqrt(uz) / dist
                scale = U.Hcoef;
            } else {
                const double dist = sqrt(dx * dx + dy * dy + dz * dz);
                ux = dx / dist; uy = dy / dist; uz = dz / dist;
                a = U.Hcoef * sqrt(uz) / dist;                                      // :213-219
                scale = 1.0;
            }
            const double px = L.source_pol == 0, py = L.source_pol == 1, pz = L.source_pol == 2;
            dHx = (uy * pz - uz * py) * a;
            dHy = (uz * px - ux * pz) * a;
            const double dHz = (ux * py - uy * px) * a;
            dEx = (dHy * uz - dHz * uy) * L.Z0;                                     // :221
            dEy = (dHz * ux - dHx * uz) * L.Z0;                                     // :222
        }
        Acc Ex = {0, 0}, Ey = {0, 0}, Hx = {0, 0}, Hy = {0, 0};
        if (ring >= 0) {
            // ---------------- periphery (:148-354)
            const RingAux &ra = reinterpret_cast<const RingAux *>(L.ring_aux)[ring];
            const RingAuxF &rf = reinterpret_cast<const RingAuxF *>(L.ring_aux_f32)[ring];
            const int gc = ra.gc;
            const double apg = ra.apg, rc = ra.rc;                                  // :161
            // grating copy: rot = round(phi / angle_per_grating) * angle_per_grating (:167, half to even)
            double kd;
            bool have_k = false;
            if (FAST && !FIX) {
                // fp32 screen: |error of phi_f / apg| < guard (atan2f <= 2 ulp, float inputs), so unless the
                // quotient is within `guard` of a half-integer the rounded index equals the float64 one
                const float yf = (float)y;
                const float kf = atan2f(yf, (float)x) * rf.inv_apg;
                const float kr = rintf(kf);
                if (fabsf(kf - kr) < 0.5f - rf.guard && fabsf(yf) > 1e-30f) { kd = (double)kr; have_k = true; }
            }
            if (FIX) {
                kd = (double)forced_cell;               // the copy index the reference's own arithmetic gives (fix-up launch)
            } else if (!have_k) {
                const double q = atan2(y, x) / apg;
                kd = rint(q);
                // a sample (numerically) ON the boundary between two grating copies: round() hinges on the last bit of
                // atan2, which differs between math libraries -- reported like an exact nearest-cell tie
                if (out.tie_count && fabs(fabs(q - kd) - 0.5) < 1e-9) {
                    const int slot = atomicAdd(out.tie_count, 1);
                    if (slot < out.tie_capacity) out.tie_list[slot] = i * L.ny + j;
                }
            }
            const double rot = kd * apg;
            double s, c;
            if (FAST) sincospi(kd * (apg * 0.31830988618379067154), &s, &c);       // rot / pi: no large-argument reduction
            else sincos(rot, &s, &c);
            const double uxp = ux * c + uy * s, uyp = -ux * s + uy * c;             // :195-196
            const double xp = x * c + y * s - rc, yp = -x * s + y * c;              // :200-201
            const double Hxp_w = dHx * c + dHy * s, Hyp_w = -dHx * s + dHy * c;     // :231-234
            Acc Exp = {0, 0}, Eyp = {0, 0}, Hxp = {0, 0}, Hyp = {0, 0};
            if (gc >= 0 && gc < L.n_packs) {
                Interp3 q;
                q.i2 = ra.i2; q.t2 = ra.t2;
                // weights: H_xp_weight = Hyp, H_yp_weight = Hxp (:246-247)
                // float32 slice of the ring: the tables of its collection interpolated at the ring's grating period
                const mlb_table_pack &pk = L.packs[gc];
                const float4 *rt = reinterpret_cast<const float4 *>(L.ring_tables) + (size_t)ring * (L.ring_table_stride >> 2);
                order_loop<STATS, FAST, Acc, W>(pk, U, L.Z0, uxp, uyp, ra.gp, true, true, q, xp, yp, ra.qx, ra.qy,
                                                rf.fqx, rf.fqy, rf.inv_fqx, rf.inv_fqy, (W)Hyp_w, (W)Hxp_w, out,
                                                Exp, Eyp, Hxp, Hyp, rt, pk.n_uy * 2, 2, pk.n_ux * pk.n_uy * 2);
            }
            if (!L.plane_wave) {                                                    // :337-346
                const double gx = rc * c, gy = rc * s;                              // :170-171
                const double path2 = (gx - L.source_x) * (gx - L.source_x) + (gy - L.source_y) * (gy - L.source_y) +
                                     L.source_z * L.source_z;
                const double path = FAST ? path2 * rsqrt_fast(path2) : sqrt(path2);
                if constexpr (FAST) {
                    const cf e = expi_fast(U.kvac * path);
                    Exp = Exp * e; Eyp = Eyp * e; Hxp = Hxp * e; Hyp = Hyp * e;
                } else {
                    const cplx e = expi(U.kvac * path);
                    Exp = Exp * e; Eyp = Eyp * e; Hxp = Hxp * e; Hyp = Hyp * e;
                }
            }
            const W cw = (W)c, sw = (W)s;
            Ex = {Exp.re * cw - Eyp.re * sw, Exp.im * cw - Eyp.im * sw};           // :351-354
            Ey = {Exp.re * sw + Eyp.re * cw, Exp.im * sw + Eyp.im * cw};
            Hx = {Hxp.re * cw - Hyp.re * sw, Hxp.im * cw - Hyp.im * sw};
            Hy = {Hxp.re * sw + Hyp.re * cw, Hxp.im * sw + Hyp.im * cw};
            local_power = (dEx * dHy - dEy * dHx) * (scale * scale);                // :474
        } else if (in_center) {
            // ---------------- centre (:359-466): nearest cell through the bin grid
            local_power = (dEx * dHy - dEy * dHx) * (scale * scale);
            int best = -1, best_orig = -1;
            double best_d2 = CUDART_INF;
            bool tie = false;
            if (FIX) {
                best = forced_cell;
            } else if (L.n_cells > 0) {
                const double fx = (x - L.bin_x0) / L.bin_size, fy = (y - L.bin_y0) / L.bin_size;
                int bx = (int)floor(fx), by = (int)floor(fy);
                bx = min(max(bx, 0), L.nbx - 1);
                by = min(max(by, 0), L.nby - 1);
                // distance from the sample to the border of its own bin: after ring k of bins every unvisited cell is at
                // least (k + that) bins away (0 if the sample lies outside the bin grid)
                const double inside = fmin(fmin(fx - bx, bx + 1 - fx), fmin(fy - by, by + 1 - fy));
                const double edge = fmax(inside, 0.0) * L.bin_size;
                const int kmax = max(L.nbx, L.nby);
                for (int k = 0; k <= kmax; ++k) {
                    for (int iy = by - k; iy <= by + k; ++iy) {
                        if (iy < 0 || iy >= L.nby) continue;
                        const bool edge_row = (iy == by - k) || (iy == by + k);
                        const int step = edge_row ? 1 : 2 * k;          // interior rows: only the two end bins
                        for (int ix = bx - k; ix <= bx + k; ix += (step > 0 ? step : 1)) {
                            if (ix < 0 || ix >= L.nbx) continue;
                            const int bb = iy * L.nbx + ix;
                            for (int cidx = L.bin_start[bb]; cidx < L.bin_start[bb + 1]; ++cidx) {
                                const double ddx = L.cell_x[cidx] - x, ddy = L.cell_y[cidx] - y;
                                const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
                                if (d2 < best_d2) {
                                    best_d2 = d2; best = cidx; best_orig = L.cell_orig[cidx]; tie = false;
                                } else if (d2 == best_d2) {                     // exact tie: highest row wins, and is reported
                                    tie = true;
                                    const int orig = L.cell_orig[cidx];
                                    if (orig > best_orig) { best = cidx; best_orig = orig; }
                                }
                            }
                        }
                    }
                    const double reach = (k * L.bin_size + edge) * (1.0 - 1e-12);
                    if (best >= 0 && best_d2 < reach * reach) break;
                }
                if (tie && out.tie_count) {
                    const int slot = atomicAdd(out.tie_count, 1);
                    if (slot < out.tie_capacity) out.tie_list[slot] = i * L.ny + j;
                }
            }
            if (best >= 0) {
                const double cx = L.cell_x[best], cy = L.cell_y[best];
                const double which = (double)L.cell_which[best];                    // :367
                const double qx = 2.0 * PI / L.hex_x_period, qy = 2.0 * PI / L.hex_y_period;
                const float fqx = (float)(qx * U.inv_kvac), fqy = (float)(qy * U.inv_kvac);
                Interp3 q;
                q.i2 = 0; q.t2 = 0.0;
                // weights un-rotated: H_x_weight = Hy, H_y_weight = Hx (:375-376)
                // float32 slice of the cell kind: an exact node of the third axis (:411), so bilinear in (ux, uy)
                const int node = min(max(L.cell_which[best], 0), L.hex.n_g - 1);
                const float4 *ht = reinterpret_cast<const float4 *>(L.hex.values_f32) + node * 2;
                order_loop<STATS, FAST, Acc, W>(L.hex, U, L.Z0, ux, uy, which, false, false, q, x - cx, y - cy, qx, qy, fqx, fqy,
                                                1.0f / fqx, 1.0f / fqy, (W)dHy, (W)dHx, out, Ex, Ey, Hx, Hy,
                                                ht, L.hex.n_uy * L.hex.n_g * 2, L.hex.n_g * 2, L.hex.n_ux * L.hex.n_uy * L.hex.n_g * 2);
                if (!L.plane_wave) {                                                // :453-461
                    const double path2 = (cx - L.source_x) * (cx - L.source_x) + (cy - L.source_y) * (cy - L.source_y) +
                                         L.source_z * L.source_z;
                    const double path = FAST ? path2 * rsqrt_fast(path2) : sqrt(path2);
                    if constexpr (FAST) {
                        const cf e = expi_fast(U.kvac * path);
                        Ex = Ex * e; Ey = Ey * e; Hx = Hx * e; Hy = Hy * e;
                    } else {
                        const cplx e = expi(U.kvac * path);
                        Ex = Ex * e; Ey = Ey * e; Hx = Hx * e; Hy = Hy * e;
                    }
                }
            }
        }
        const size_t off = (size_t)i * out.ld + j;
        if constexpr (FAST) {
            // the dipole scale left out of the fp32 weights, applied in float64
            reinterpret_cast<float2 *>(out.F[0])[off] = make_float2((float)(Ex.re * scale), (float)(Ex.im * scale));
            reinterpret_cast<float2 *>(out.F[1])[off] = make_float2((float)(Ey.re * scale), (float)(Ey.im * scale));
            reinterpret_cast<float2 *>(out.F[2])[off] = make_float2((float)(Hx.re * scale), (float)(Hx.im * scale));
            reinterpret_cast<float2 *>(out.F[3])[off] = make_float2((float)(Hy.re * scale), (float)(Hy.im * scale));
        } else {
            store_c(out.F[0], off, Ex, out.out_is_double);
            store_c(out.F[1], off, Ey, out.out_is_double);
            store_c(out.F[2], off, Hx, out.out_is_double);
            store_c(out.F[3], off, Hy, out.out_is_double);
        }
    }
    // incident power through the lens (:474-477), deterministic per-warp partial sums (no block barrier: the
    // warps of a block finish at very different times)
    if (FIX) return;
    double v = local_power;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0)
        out.power_warp_sums[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (NF_THREADS / 32) + (threadIdx.x >> 5)] = v;
}

// Per-lens derived data (mlb_nearfield_prepare): ring records and the ring bin table.
__global__ void nearfield_prepare_kernel(const __grid_constant__ mlb_lens_desc L, double kvac) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double PI = 3.14159265358979323846;
    if (t < L.n_rings) {
        RingAux ra;
        ra.rc = L.r_center[t];
        ra.gp = L.grating_period[t];
        ra.apg = 2.0 * PI / L.num_around[t];                                       // :161
        ra.lat = ra.rc * ra.apg;                                                   // :165
        ra.qx = 2.0 * PI / ra.gp;
        ra.qy = 2.0 * PI / ra.lat;
        ra.gc = L.gc_index[t];
        ra.i2 = 0; ra.t2 = 0.0;
        if (ra.gc >= 0 && ra.gc < L.n_packs) {
            Interp3 q;
            locate3(L.packs[ra.gc], ra.gp, q);
            ra.i2 = q.i2; ra.t2 = q.t2;
        }
        reinterpret_cast<RingAux *>(L.ring_aux)[t] = ra;
        RingAuxF rf;
        rf.fqx = (float)(ra.qx / kvac); rf.fqy = (float)(ra.qy / kvac);
        rf.inv_fqx = 1.0f / rf.fqx; rf.inv_fqy = 1.0f / rf.fqy;
        rf.inv_apg = (float)(1.0 / ra.apg);
        // |error of atan2f(yf, xf) * inv_apg| <= (2 ulp of pi-sized results 4.8e-7 + input rounding 6e-8 + the roundings
        // of inv_apg and of the product, 1.2e-7 * pi) * inv_apg = 9.2e-7 * inv_apg; guard = 1.6 x that
        rf.guard = fminf(0.5f, 1.5e-6f * rf.inv_apg + 1e-6f);
        rf.pad0 = rf.pad1 = 0.f;
        reinterpret_cast<RingAuxF *>(L.ring_aux_f32)[t] = rf;
    }
    if (t <= L.n_lut) {
        int cnt = L.n_rings + 1;
        if (t < L.n_lut) {
            const double edge = t * (L.lut_r_max / L.n_lut);
            int lo = 0, hi = L.n_rings + 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (L.ring_boundary[mid] < edge) lo = mid + 1; else hi = mid;
            }
            cnt = lo;
        }
        L.ring_lut[t] = cnt;
    }
}

// Per-ring float32 slices (mlb_nearfield_prepare): ring r's collection tables interpolated at the ring's grating period,
// slice[o][iu][iv][slot] = (1 - t2) T[o][iu][iv][i2][slot] + t2 T[o][iu][iv][i2 + 1][slot]  (scipy RGI's linear rule along
// the third axis, formed in float64, rounded once).  The complex64-output kernel then interpolates bilinearly.
__global__ void nearfield_ring_tables_kernel(const __grid_constant__ mlb_lens_desc L) {
    const int ring = blockIdx.y;
    const RingAux &ra = reinterpret_cast<const RingAux *>(L.ring_aux)[ring];
    if (ra.gc < 0 || ra.gc >= L.n_packs) return;
    const mlb_table_pack &p = L.packs[ra.gc];
    const int n = p.n_orders * p.n_ux * p.n_uy * 4;                  // complex entries of the slice
    float2 *dst = reinterpret_cast<float2 *>(L.ring_tables + (size_t)ring * L.ring_table_stride);
    const double2 *src = reinterpret_cast<const double2 *>(p.values);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int slot = e & 3, cellidx = e >> 2;                    // cellidx = (o * n_ux + iu) * n_uy + iv
        const double2 a = src[((size_t)cellidx * p.n_g + ra.i2) * 4 + slot];
        const double2 b = src[((size_t)cellidx * p.n_g + ra.i2 + 1) * 4 + slot];
        const double w0 = 1.0 - ra.t2, w1 = ra.t2;
        dst[e] = make_float2((float)(a.x * w0 + b.x * w1), (float)(a.y * w0 + b.y * w1));
    }
}

// mlb_table_pack_build: the interpolator value arrays as the reference holds them -- one (n_ux, n_uy, n_g) complex128 array per
// (order, slot), stacked [order][slot][iu][iv][ig] -- into the layout the assembly kernel gathers from,
// [order][iu][iv][ig][slot] (the four values a corner needs adjacent), as complex128 and as complex64.
__global__ void table_pack_kernel(const double2 *__restrict__ raw, int n_orders, int cells, double2 *__restrict__ values,
                                  float2 *__restrict__ values_f32) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // output index
    const long long total = (long long)n_orders * cells * 4;
    if (e >= total) return;
    const int slot = (int)(e & 3);
    const long long oc = e >> 2;                                               // order * cells + cell
    const long long o = oc / cells, cell = oc - o * cells;
    const double2 v = raw[(o * 4 + slot) * cells + cell];
    values[e] = v;
    values_f32[e] = make_float2((float)v.x, (float)v.y);
}

__global__ void table_eval_kernel(const double *__restrict__ axes, int n0, int n1, int n2,
                                  const double2 *__restrict__ values, const double *__restrict__ pts, int n,
                                  double2 *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double *a0 = axes, *a1 = axes + n0, *a2 = axes + n0 + n1;
    const double u0 = pts[3 * t], u1 = pts[3 * t + 1], u2 = pts[3 * t + 2];
    const int i0 = find_interval(a0, n0, u0), i1 = find_interval(a1, n1, u1), i2 = find_interval(a2, n2, u2);
    const double t0 = (u0 - a0[i0]) / (a0[i0 + 1] - a0[i0]);
    const double t1 = (u1 - a1[i1]) / (a1[i1 + 1] - a1[i1]);
    const double t2 = (u2 - a2[i2]) / (a2[i2 + 1] - a2[i2]);
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double w = (a ? t0 : 1.0 - t0) * (b ? t1 : 1.0 - t1) * (c ? t2 : 1.0 - t2);
                const double2 v = values[((size_t)(i0 + a) * n1 + (i1 + b)) * n2 + (i2 + c)];
                re += v.x * w;
                im += v.y * w;
            }
    out[t] = make_double2(re, im);
}

}  // namespace mlb

static int g_nf_lg_wy = 3;       // warp tile: 2^lg_wy samples along y times 32 / 2^lg_wy along x (mlb_nearfield_tune(100 + lg_wy)); 8 x 4 measured fastest
static int g_nf_minblocks = 8;   // 64 registers, 8 blocks/SM: fastest on B200 since the round-2 rework (scripts/tune_nearfield.py)
/* tuning knob: minimum resident blocks per SM the complex64 kernel is compiled for (1, 5, 6 or 8) */
extern "C" int mlb_nearfield_tune(int min_blocks) {
    if (min_blocks >= 102 && min_blocks <= 105) { g_nf_lg_wy = min_blocks - 100; return MLB_OK; }   // warp tile shape
    MLB_REQUIRE(min_blocks == 1 || min_blocks == 5 || min_blocks == 6 || min_blocks == 8,
                "mlb_nearfield_tune: min_blocks must be 1, 5, 6 or 8 (or 102..105 for the warp tile)");
    g_nf_minblocks = min_blocks;
    return MLB_OK;
}

static dim3 nf_grid(int nx, int ny) {
    const int wy = 1 << g_nf_lg_wy, wx = 32 >> g_nf_lg_wy, by = wy * (mlb::NF_THREADS / 32);
    return dim3((ny + by - 1) / by, (nx + wx - 1) / wx);
}
/* one incident-power partial sum per warp */
extern "C" int mlb_nearfield_blocks(int nx, int ny) {
    const dim3 g = nf_grid(nx, ny);
    return (int)(g.x * g.y) * (mlb::NF_THREADS / 32);
}

static int check_pack(const mlb_table_pack &p, const char *what) {
    MLB_REQUIRE(p.axes && p.values && p.values_f32 && p.orders, "mlb_nearfield_assemble: %s pack has NULL arrays", what);
    MLB_REQUIRE(mlb::aligned16(p.values_f32), "mlb_nearfield_assemble: %s pack values_f32 not 16-byte aligned", what);
    MLB_REQUIRE(p.order_map && p.order_radius >= 0 && p.order_radius <= 64, "mlb_nearfield_assemble: %s pack has no order map", what);
    MLB_REQUIRE(p.n_ux >= 2 && p.n_uy >= 2 && p.n_g >= 2 && p.n_orders >= 0,
                "mlb_nearfield_assemble: %s pack needs >= 2 nodes per axis (%d,%d,%d)", what, p.n_ux, p.n_uy, p.n_g);
    MLB_REQUIRE(mlb::aligned16(p.values), "mlb_nearfield_assemble: %s pack values not 16-byte aligned", what);
    return MLB_OK;
}

static int check_desc(const mlb_lens_desc &L, const char *who) {
    MLB_REQUIRE(L.n_rings >= 1 && L.ring_boundary && L.r_center && L.grating_period && L.num_around && L.gc_index,
                "%s: the periphery needs >= 1 ring (nearfield.py:87-93 dereferences it)", who);
    MLB_REQUIRE(L.n_packs >= 1 && L.n_packs <= MLB_MAX_PACKS, "%s: %d collections (max %d)", who, L.n_packs, MLB_MAX_PACKS);
    for (int g = 0; g < L.n_packs; ++g)
        if (int rc = check_pack(L.packs[g], "collection")) return rc;
    MLB_REQUIRE(L.ring_aux && L.ring_aux_f32 && L.ring_lut && L.n_lut >= 1 && L.lut_r_max > 0,
                "%s: ring_aux / ring_aux_f32 / ring_lut buffers missing (see mlb_nearfield_prepare)", who);
    MLB_REQUIRE(mlb::aligned16(L.ring_aux) && mlb::aligned16(L.ring_aux_f32), "%s: ring_aux buffers not 16-byte aligned", who);
    MLB_REQUIRE(L.wavelength > 0 && L.n_glass > 0, "%s: bad wavelength / n_glass", who);
    MLB_REQUIRE(L.ring_tables && L.ring_table_stride > 0, "%s: ring_tables missing (see mlb_nearfield_prepare)", who);
    return MLB_OK;
}

extern "C" long long mlb_nearfield_ring_table_floats(const mlb_lens_desc *h_desc) {
    if (!h_desc) return 0;
    long long need = 4;
    for (int g = 0; g < h_desc->n_packs && g < MLB_MAX_PACKS; ++g) {
        const long long v = 8LL * h_desc->packs[g].n_orders * h_desc->packs[g].n_ux * h_desc->packs[g].n_uy;
        if (v > need) need = v;
    }
    return need;
}

extern "C" int mlb_nearfield_prepare(const mlb_lens_desc *h_desc, void *stream) {
    MLB_REQUIRE(h_desc, "mlb_nearfield_prepare: NULL descriptor");
    const mlb_lens_desc &L = *h_desc;
    if (int rc = check_desc(L, "mlb_nearfield_prepare")) return rc;
    const double kvac = 2.0 * 3.14159265358979323846 / L.wavelength;
    const int n = (L.n_rings > L.n_lut + 1) ? L.n_rings : L.n_lut + 1;
    mlb::nearfield_prepare_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(L, kvac);
    if (int rc = mlb::check_launch("mlb_nearfield_prepare")) return rc;
    long long need = 0;
    for (int g = 0; g < L.n_packs; ++g) {
        const long long v = 8LL * L.packs[g].n_orders * L.packs[g].n_ux * L.packs[g].n_uy;
        if (v > need) need = v;
    }
    MLB_REQUIRE(L.ring_tables && L.ring_table_stride >= need && L.ring_table_stride % 4 == 0 && mlb::aligned16(L.ring_tables),
                "mlb_nearfield_prepare: ring_tables needs n_rings x %lld floats (see mlb_nearfield_ring_table_floats)", need);
    const int per = (int)((need / 2 + 255) / 256);
    mlb::nearfield_ring_tables_kernel<<<dim3(per > 16 ? 16 : (per < 1 ? 1 : per), L.n_rings), 256, 0, (cudaStream_t)stream>>>(L);
    return mlb::check_launch("mlb_nearfield_prepare(ring tables)");
}

static int nearfield_launch(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld, int out_is_double,
                            double *power_block_sums, long long *stats, int want_stats, int *violation, int *tie_count,
                            int *tie_list, int tie_capacity, const int *fix_samples, const int *fix_cells, int n_fix,
                            void *stream, const char *who) {
    const bool fix = fix_samples != nullptr;
    MLB_REQUIRE(h_desc && Ex && Ey && Hx && Hy && (fix || power_block_sums) && violation, "%s: NULL pointer", who);
    const mlb_lens_desc &L = *h_desc;
    MLB_REQUIRE(L.nx > 0 && L.ny > 0 && ld >= L.ny, "%s: bad grid (%d,%d,ld=%d)", who, L.nx, L.ny, ld);
    MLB_REQUIRE(L.x_pts && L.y_pts, "%s: NULL sample coordinates", who);
    if (int rc = check_desc(L, who)) return rc;
    if (L.n_cells > 0) {
        if (int rc = check_pack(L.hex, "hexgridset")) return rc;
        MLB_REQUIRE(L.cell_x && L.cell_y && L.cell_which && L.cell_orig && L.bin_start && L.nbx > 0 && L.nby > 0 &&
                        L.bin_size > 0,
                    "%s: bad centre-cell bin grid", who);
    }
    MLB_REQUIRE(L.plane_wave || L.source_z < 0, "%s: source_z must be negative (nearfield.py:84)", who);
    MLB_REQUIRE(L.source_pol >= 0 && L.source_pol <= 2 && !(L.plane_wave && L.source_pol == 2),
                "%s: bad source polarisation (nearfield.py:85, :224)", who);
    MLB_REQUIRE(!want_stats || stats, "%s: want_stats needs a stats buffer", who);
    MLB_REQUIRE(L.ny <= (1 << 24) && L.nx <= 65535 && (long long)L.nx * L.ny < (1LL << 31), "%s: grid too large", who);
    MLB_REQUIRE(!tie_count || (tie_list && tie_capacity > 0), "%s: tie_count needs a tie_list", who);
    MLB_REQUIRE(!fix || (fix_cells && n_fix > 0 && L.n_cells > 0), "%s: bad fix-up list", who);
    mlb::NfOut out;
    out.F[0] = Ex; out.F[1] = Ey; out.F[2] = Hx; out.F[3] = Hy;
    out.power_warp_sums = power_block_sums; out.stats = stats; out.violation = violation;
    out.ld = ld; out.out_is_double = out_is_double; out.lg_wy = g_nf_lg_wy;
    out.tie_count = tie_count; out.tie_list = tie_list; out.tie_capacity = tie_capacity;
    out.fix_samples = fix_samples; out.fix_cells = fix_cells; out.n_fix = n_fix;
    // launch-uniform scalars in IEEE float64, the expressions of nearfield.py:213 / :262 / :287
    const double PI = 3.14159265358979323846;
    mlb::NfUniform U;
    U.kvac = 2.0 * PI / L.wavelength;
    U.kg = 2.0 * PI * L.n_glass / L.wavelength;
    U.inv_kvac = 1.0 / U.kvac;
    U.inv_kg_n = 1.0 / (U.kg * L.n_glass);
    U.Hcoef = L.c0 * (U.kvac * U.kvac) * L.dipole_moment / (4.0 * PI);
    U.lut_scale = L.n_lut / L.lut_r_max;
    U.ng2 = (float)((U.kg / U.kvac) * (U.kg / U.kvac));
    U.cf = (float)(L.Z0 * U.inv_kg_n * U.kvac);
    const cudaStream_t st = (cudaStream_t)stream;
    if (fix) {
        const dim3 grid((n_fix + mlb::NF_THREADS - 1) / mlb::NF_THREADS);
        if (out_is_double) mlb::nearfield_kernel<false, false, 1, true><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
        else mlb::nearfield_kernel<false, true, 1, true><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
        return mlb::check_launch(who);
    }
    const dim3 grid = nf_grid(L.nx, L.ny);
    if (want_stats) {
        if (out_is_double) mlb::nearfield_kernel<true, false, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
        else mlb::nearfield_kernel<true, true, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
    } else {
        if (out_is_double) mlb::nearfield_kernel<false, false, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
        else if (g_nf_minblocks == 5) mlb::nearfield_kernel<false, true, 5><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
        else if (g_nf_minblocks == 6) mlb::nearfield_kernel<false, true, 6><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
        else if (g_nf_minblocks == 8) mlb::nearfield_kernel<false, true, 8><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
        else mlb::nearfield_kernel<false, true, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, U, out);
    }
    return mlb::check_launch(who);
}

extern "C" int mlb_nearfield_assemble(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld,
                                      int out_is_double, double *power_block_sums, long long *stats, int want_stats,
                                      int *violation, void *stream) {
    return nearfield_launch(h_desc, Ex, Ey, Hx, Hy, ld, out_is_double, power_block_sums, stats, want_stats, violation,
                            nullptr, nullptr, 0, nullptr, nullptr, 0, stream, "mlb_nearfield_assemble");
}

extern "C" int mlb_nearfield_assemble_ties(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld,
                                           int out_is_double, double *power_block_sums, long long *stats, int want_stats,
                                           int *violation, int *tie_count, int *tie_list, int tie_capacity, void *stream) {
    MLB_REQUIRE(tie_count && tie_list && tie_capacity > 0, "mlb_nearfield_assemble_ties: tie buffers missing");
    return nearfield_launch(h_desc, Ex, Ey, Hx, Hy, ld, out_is_double, power_block_sums, stats, want_stats, violation,
                            tie_count, tie_list, tie_capacity, nullptr, nullptr, 0, stream, "mlb_nearfield_assemble_ties");
}

extern "C" int mlb_nearfield_fixup(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld,
                                   int out_is_double, const int *fix_samples, const int *fix_cells, int n_fix,
                                   int *violation, void *stream) {
    MLB_REQUIRE(fix_samples && fix_cells && n_fix > 0, "mlb_nearfield_fixup: empty fix-up list");
    return nearfield_launch(h_desc, Ex, Ey, Hx, Hy, ld, out_is_double, nullptr, nullptr, 0, violation, nullptr, nullptr, 0,
                            fix_samples, fix_cells, n_fix, stream, "mlb_nearfield_fixup");
}

extern "C" int mlb_table_pack_build(const double *raw, int n_orders, int n_ux, int n_uy, int n_g, double *values,
                              float *values_f32, void *stream) {
    MLB_REQUIRE(raw && values && values_f32 && n_orders >= 0 && n_ux >= 2 && n_uy >= 2 && n_g >= 2,
                "mlb_table_pack_build: bad arguments");
    MLB_REQUIRE(mlb::aligned16(raw) && mlb::aligned16(values) && mlb::aligned16(values_f32), "mlb_table_pack_build: buffers not 16-byte aligned");
    if (n_orders == 0) return MLB_OK;
    const int cells = n_ux * n_uy * n_g;
    const long long total = (long long)n_orders * cells * 4;
    mlb::table_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2 *>(raw), n_orders, cells, reinterpret_cast<double2 *>(values),
        reinterpret_cast<float2 *>(values_f32));
    return mlb::check_launch("mlb_table_pack_build");
}

extern "C" int mlb_table_eval(const double *axes, int n0, int n1, int n2, const double *values, const double *pts,
                              int n, double *out, void *stream) {
    MLB_REQUIRE(axes && values && (n == 0 || (pts && out)), "mlb_table_eval: NULL pointer");
    MLB_REQUIRE(n0 >= 2 && n1 >= 2 && n2 >= 2 && n >= 0, "mlb_table_eval: bad sizes");
    if (n == 0) return MLB_OK;
    mlb::table_eval_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        axes, n0, n1, n2, reinterpret_cast<const double2 *>(values), pts, n, reinterpret_cast<double2 *>(out));
    return mlb::check_launch("mlb_table_eval");
}
