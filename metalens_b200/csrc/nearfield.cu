// Aperture-field assembly: fused build_nearfield (reference nearfield.py:66-480).
//
// One thread per aperture sample (i = x index, j = y index, j fastest so the complex output rows
// are written coalesced).  Everything the reference does with ~40 full-array numpy passes, four
// scipy RegularGridInterpolator calls per diffraction order and a cKDTree query happens here in
// registers, in float64 (geometry and phases need it; B200 has a full-rate FP64 pipe
// for this amount of work).  Tables are a few hundred KB per collection and stay L1/L2 resident;
// neighbouring threads hit the same interpolation cell, so the gathers are mostly broadcasts.
#include <math_constants.h>

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

__device__ __forceinline__ Interp3 locate(const mlb_table_pack &p, double u0, double u1, double u2) {
    Interp3 q;
    const double *a0 = p.axes, *a1 = p.axes + p.n_ux, *a2 = p.axes + p.n_ux + p.n_uy;
    q.i0 = find_interval(a0, p.n_ux, u0);
    q.i1 = find_interval(a1, p.n_uy, u1);
    q.i2 = find_interval(a2, p.n_g, u2);
    q.t0 = (u0 - a0[q.i0]) / (a0[q.i0 + 1] - a0[q.i0]);
    q.t1 = (u1 - a1[q.i1]) / (a1[q.i1 + 1] - a1[q.i1]);
    q.t2 = (u2 - a2[q.i2]) / (a2[q.i2 + 1] - a2[q.i2]);
    return q;
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
// to [-pi, pi] in float64 (exact to ~1e-16 turns) and the sine/cosine taken in fp32 (abs. error
// ~1e-7, below the output's own rounding).
template <bool FAST>
__device__ __forceinline__ cplx expi(double x) {
    if (FAST) {
        const double t = x * 0.15915494309189535;                 // x / 2 pi
        const float f = (float)((t - rint(t)) * 6.283185307179586);
        float s, c;
        sincosf(f, &s, &c);
        return {(double)c, (double)s};
    }
    double s, c;
    sincos(x, &s, &c);
    return {c, s};
}

// fp32 versions for the complex64 output path: tables read from the float2 copy of the pack, the
// order's contribution formed in fp32 (relative error ~1e-7, below the output rounding) and added
// to the float64 per-sample accumulators.
struct cf { float re, im; };
__device__ __forceinline__ cf operator+(cf a, cf b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cf operator*(cf a, cf b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cf operator*(cf a, float s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ cf operator*(float s, cf a) { return {a.re * s, a.im * s}; }

__device__ __forceinline__ void gather4f(const mlb_table_pack &p, int order, const Interp3 &q, cf (&amp)[4]) {
#pragma unroll
    for (int s = 0; s < 4; ++s) amp[s] = {0.f, 0.f};
    const size_t per_order = (size_t)p.n_ux * p.n_uy * p.n_g;
    const float4 *__restrict__ base = reinterpret_cast<const float4 *>(p.values_f32) + (size_t)order * per_order * 2;
    const float t0 = (float)q.t0, t1 = (float)q.t1, t2 = (float)q.t2;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const float wa = a ? t0 : 1.f - t0;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const float wb = b ? t1 : 1.f - t1;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float w = wa * wb * (c ? t2 : 1.f - t2);
                const float4 *v = base + (((size_t)(q.i0 + a) * p.n_uy + (q.i1 + b)) * p.n_g + (q.i2 + c)) * 2;
                const float4 x0 = __ldg(v), x1 = __ldg(v + 1);          // slots (0,1) and (2,3)
                amp[0].re = fmaf(x0.x, w, amp[0].re); amp[0].im = fmaf(x0.y, w, amp[0].im);
                amp[1].re = fmaf(x0.z, w, amp[1].re); amp[1].im = fmaf(x0.w, w, amp[1].im);
                amp[2].re = fmaf(x1.x, w, amp[2].re); amp[2].im = fmaf(x1.y, w, amp[2].im);
                amp[3].re = fmaf(x1.z, w, amp[3].re); amp[3].im = fmaf(x1.w, w, amp[3].im);
            }
        }
    }
}

__device__ __forceinline__ void add_order_f(const cf (&amp)[4], float Hw_x, float Hw_y, float kx, float ky, float kz,
                                            float f, cf phase, cplx &Ea, cplx &Eb, cplx &Ha, cplx &Hb) {
    const cf Sfy = Hw_x * amp[0] + Hw_y * amp[2];
    const cf Sfx = Hw_x * amp[1] + Hw_y * amp[3];
    const cf ea = ((Sfy * (kx * ky) + Sfx * (ky * ky + kz * kz)) * f) * phase;
    const cf eb = ((Sfy * (-kx * kx - kz * kz) + Sfx * (-kx * ky)) * f) * phase;
    const cf ha = Sfy * phase, hb = Sfx * phase;
    Ea.re += ea.re; Ea.im += ea.im; Eb.re += eb.re; Eb.im += eb.im;
    Ha.re += ha.re; Ha.im += ha.im; Hb.re += hb.re; Hb.im += hb.im;
}

__device__ __forceinline__ void add_order(const cplx (&amp)[4], double Hw_x, double Hw_y, double kx, double ky,
                                          double kz, double inv_kg_n, double Z0, cplx phase, cplx &Ea, cplx &Eb,
                                          cplx &Ha, cplx &Hb);

// one diffraction order, either precision.  kx, ky are in units of kvac in the fp32 path so that the
// products stay well inside the fp32 range; the common factor Z0 kvac / (k_g n) is applied via `f`.
template <bool FAST>
__device__ __forceinline__ void order_term(const mlb_table_pack &p, int o, const Interp3 &q, double Hw_x, double Hw_y,
                                           double kx, double ky, double kg, double kvac, double inv_kg_n, double Z0,
                                           double phase_arg, cplx &Ea, cplx &Eb, cplx &Ha, cplx &Hb) {
    if (FAST) {
        const float nx = (float)(kx / kvac), ny = (float)(ky / kvac);
        const float ng2 = (float)((kg / kvac) * (kg / kvac));
        const float nz = sqrtf(ng2 - nx * nx - ny * ny);
        const cplx ph = expi<true>(phase_arg);
        cf amp[4];
        gather4f(p, o, q, amp);
        const float f = (float)(Z0 * inv_kg_n * kvac) / nz;
        add_order_f(amp, (float)Hw_x, (float)Hw_y, nx, ny, nz, f, {(float)ph.re, (float)ph.im}, Ea, Eb, Ha, Hb);
    } else {
        const double kz = sqrt(kg * kg - kx * kx - ky * ky);
        const cplx ph = expi<false>(phase_arg);
        cplx amp[4];
        gather4(p, o, q, amp);
        add_order(amp, Hw_x, Hw_y, kx, ky, kz, inv_kg_n, Z0, ph, Ea, Eb, Ha, Hb);
    }
}

// one diffraction order's contribution (nearfield.py:306-327 / :420-441), both incident
// polarisations and both amplitudes at once:
//   E_a += Z0 [ S_fy kx ky + S_fx (ky^2+kz^2) ] / (k_g kz n) * phase
//   E_b += Z0 [ S_fy (-kx^2-kz^2) - S_fx kx ky ] / (k_g kz n) * phase
//   H_a += S_fy * phase ;  H_b += S_fx * phase        with S_f* = Hw_x a_x,f* + Hw_y a_y,f*
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

template <bool STATS>
__device__ __forceinline__ void record(const mlb_table_pack &p, int order, double u0, double u1, double u2, bool check2,
                                       long long *stats, int *violation) {
    const bool bad = (u0 < p.bounds[0]) | (u0 > p.bounds[1]) | (u1 < p.bounds[2]) | (u1 > p.bounds[3]) |
                     (check2 & ((u2 < p.bounds[4]) | (u2 > p.bounds[5])));
    if (bad) atomicOr(violation, 1);
    if (STATS) {
        long long *s = stats + (size_t)(p.stats_slot + order) * MLB_STATS_PER_ORDER;
        atomicAdd(reinterpret_cast<unsigned long long *>(s), 1ULL);
        atomicMin(s + 1, enc_f64(u0)); atomicMax(s + 2, enc_f64(u0));
        atomicMin(s + 3, enc_f64(u1)); atomicMax(s + 4, enc_f64(u1));
        atomicMin(s + 5, enc_f64(u2)); atomicMax(s + 6, enc_f64(u2));
    }
}

constexpr int NF_THREADS = 128;

struct NfOut {
    void *F[4];
    double *power_block_sums;
    long long *stats;
    int *violation;
    int ld, out_is_double;
};

__device__ __forceinline__ void store_c(void *base, size_t off, cplx v, int is_double) {
    if (is_double) reinterpret_cast<double2 *>(base)[off] = make_double2(v.re, v.im);
    else reinterpret_cast<float2 *>(base)[off] = make_float2((float)v.re, (float)v.im);
}

template <bool STATS, bool FAST, int MINB>
__global__ void __launch_bounds__(NF_THREADS, MINB) nearfield_kernel(const __grid_constant__ mlb_lens_desc L, const NfOut out) {
    const int j = blockIdx.x * NF_THREADS + threadIdx.x;   // y index (fast)
    const int i = blockIdx.y;                              // x index
    const double PI = 3.14159265358979323846;
    double local_power = 0.0;
    if (j < L.ny) {
        const double x = L.x_pts[i], y = L.y_pts[j];
        const double r = sqrt(x * x + y * y);                                  // nearfield.py:118
        // which_ring = searchsorted(boundaries, r) - 1  (left-biased)          :125-128
        int lo = 0, hi = L.n_rings + 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (L.ring_boundary[mid] < r) lo = mid + 1; else hi = mid;
        }
        int ring = lo - 1;
        const bool in_center = (ring == -1);
        if (ring == L.n_rings) ring = -1;

        const double kvac = 2.0 * PI / L.wavelength, kg = 2.0 * PI * L.n_glass / L.wavelength;
        const double inv_kg_n = 1.0 / (kg * L.n_glass);
        // incident direction and dipole field (:172-228)
        double ux = 0.0, uy = 0.0, uz = 1.0, dHx, dHy, dEx, dEy;
        if (L.plane_wave) {
            const double px = L.source_pol == 0 ? 1.0 : 0.0, py = L.source_pol == 1 ? 1.0 : 0.0;
            dEx = px * L.dipole_moment; dEy = py * L.dipole_moment;
            dHx = -py * L.dipole_moment / L.Z0; dHy = px * L.dipole_moment / L.Z0;
        } else {
            const double dx = x - L.source_x, dy = y - L.source_y, dz = 0.0 - L.source_z;
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);
            ux = dx / dist; uy = dy / dist; uz = dz / dist;
            const double kv = 2.0 * PI / L.wavelength;
            const double Hcoef = L.c0 * (kv * kv) * L.dipole_moment / (4.0 * PI);   // :213
            const double a = Hcoef * sqrt(uz) / dist;
            const double px = L.source_pol == 0, py = L.source_pol == 1, pz = L.source_pol == 2;
            dHx = (uy * pz - uz * py) * a;
            dHy = (uz * px - ux * pz) * a;
            const double dHz = (ux * py - uy * px) * a;
            dEx = (dHy * uz - dHz * uy) * L.Z0;                                     // :221
            dEy = (dHz * ux - dHx * uz) * L.Z0;                                     // :222
        }
        cplx Ex = {0, 0}, Ey = {0, 0}, Hx = {0, 0}, Hy = {0, 0};
        if (ring >= 0) {
            // ---------------- periphery (:148-354)
            const int gc = L.gc_index[ring];
            const double gp = L.grating_period[ring];
            const double apg = 2.0 * PI / L.num_around[ring];                       // :161
            const double rc = L.r_center[ring];
            const double lat = rc * apg;                                            // :165
            const double phi = atan2(y, x);
            const double rot = rint(phi / apg) * apg;                               // :167 (half to even)
            double s, c;
            sincos(rot, &s, &c);
            const double uxp = ux * c + uy * s, uyp = -ux * s + uy * c;             // :195-196
            const double xp = x * c + y * s - rc, yp = -x * s + y * c;              // :200-201
            const double Hxp_w = dHx * c + dHy * s, Hyp_w = -dHx * s + dHy * c;     // :231-234
            const double Hw_x = Hyp_w, Hw_y = Hxp_w;                                // :246-247
            cplx Exp = {0, 0}, Eyp = {0, 0}, Hxp = {0, 0}, Hyp = {0, 0};
            if (gc >= 0 && gc < L.n_packs) {
                const mlb_table_pack &p = L.packs[gc];
                Interp3 q;
                bool located = false;
                const double qx = 2.0 * PI / gp, qy = 2.0 * PI / lat;
                // Orders that can propagate in air satisfy |ux + ox*lambda/gp| <= 1 and |uy + oy*lambda/lat| <= 1:
                // only that (small) index box is visited, through the pack's dense (ox,oy) -> order map; an
                // fp32 screen with margin picks the box, the exact float64 test of :279 decides.
                const float fux = (float)uxp, fuy = (float)uyp, fqx = (float)(qx / kvac), fqy = (float)(qy / kvac);
                const int R = p.order_radius, W = 2 * R + 1;
                const int ox_lo = max(-R, (int)ceilf((-1.001f - fux) / fqx)), ox_hi = min(R, (int)floorf((1.001f - fux) / fqx));
                const int oy_lo = max(-R, (int)ceilf((-1.001f - fuy) / fqy)), oy_hi = min(R, (int)floorf((1.001f - fuy) / fqy));
                for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                    const double kxp = kvac * uxp + ox * qx;                               // :268
                    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
                        const int o = p.order_map[(ox + R) * W + (oy + R)];
                        if (o < 0) continue;                                               // order not in the tables
                        const double kyp = kvac * uyp + oy * qy;                           // :269
                        if (kxp * kxp + kyp * kyp <= kvac * kvac) {                        // :279
                            record<STATS>(p, o, uxp, uyp, gp, true, out.stats, out.violation);
                            if (!located) { q = locate(p, uxp, uyp, gp); located = true; }
                            // kzp (:287), phase about the grating centre (:291), table gathers, accumulation
                            order_term<FAST>(p, o, q, Hw_x, Hw_y, kxp, kyp, kg, kvac, inv_kg_n, L.Z0, kxp * xp + kyp * yp,
                                             Exp, Eyp, Hxp, Hyp);
                        }
                    }
                }
            }
            if (!L.plane_wave) {                                                    // :337-346
                const double gx = rc * c, gy = rc * s;                              // :170-171
                const double path = sqrt((gx - L.source_x) * (gx - L.source_x) + (gy - L.source_y) * (gy - L.source_y) +
                                         L.source_z * L.source_z);
                const cplx e = expi<FAST>(kvac * path);
                Exp = Exp * e; Eyp = Eyp * e; Hxp = Hxp * e; Hyp = Hyp * e;
            }
            Ex = {Exp.re * c - Eyp.re * s, Exp.im * c - Eyp.im * s};               // :351-354
            Ey = {Exp.re * s + Eyp.re * c, Exp.im * s + Eyp.im * c};
            Hx = {Hxp.re * c - Hyp.re * s, Hxp.im * c - Hyp.im * s};
            Hy = {Hxp.re * s + Hyp.re * c, Hxp.im * s + Hyp.im * c};
            local_power = dEx * dHy - dEy * dHx;                                    // :474
        } else if (in_center) {
            // ---------------- centre (:359-466): nearest cell through the bin grid
            local_power = dEx * dHy - dEy * dHx;
            int best = -1, best_orig = -1;
            double best_d2 = CUDART_INF;
            if (L.n_cells > 0) {
                int bx = (int)floor((x - L.bin_x0) / L.bin_size), by = (int)floor((y - L.bin_y0) / L.bin_size);
                bx = min(max(bx, 0), L.nbx - 1);
                by = min(max(by, 0), L.nby - 1);
                const int kmax = max(L.nbx, L.nby);
                for (int k = 0; k <= kmax; ++k) {
                    for (int iy = by - k; iy <= by + k; ++iy) {
                        if (iy < 0 || iy >= L.nby) continue;
                        const bool edge_row = (iy == by - k) || (iy == by + k);
                        const int step = edge_row ? 1 : 2 * k;          // interior rows: only the two end bins
                        for (int ix = bx - k; ix <= bx + k; ix += (step > 0 ? step : 1)) {
                            if (ix < 0 || ix >= L.nbx) continue;
                            const int b = iy * L.nbx + ix;
                            for (int cidx = L.bin_start[b]; cidx < L.bin_start[b + 1]; ++cidx) {
                                const double ddx = L.cell_x[cidx] - x, ddy = L.cell_y[cidx] - y;
                                const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
                                const int orig = L.cell_orig[cidx];
                                if (d2 < best_d2 || (d2 == best_d2 && orig > best_orig)) {   // exact ties: highest row wins
                                    best_d2 = d2; best = cidx; best_orig = orig;
                                }
                            }
                        }
                    }
                    const double reach = k * L.bin_size;
                    if (best >= 0 && best_d2 <= reach * reach) break;
                }
            }
            if (best >= 0) {
                const double cx = L.cell_x[best], cy = L.cell_y[best];
                const double which = (double)L.cell_which[best];                    // :367
                const double Hw_x = dHy, Hw_y = dHx;                                // :375-376
                const mlb_table_pack &p = L.hex;
                Interp3 q;
                bool located = false;
                const double qx = 2.0 * PI / L.hex_x_period, qy = 2.0 * PI / L.hex_y_period;
                const float fux = (float)ux, fuy = (float)uy, fqx = (float)(qx / kvac), fqy = (float)(qy / kvac);
                const int R = p.order_radius, W = 2 * R + 1;
                const int ox_lo = max(-R, (int)ceilf((-1.001f - fux) / fqx)), ox_hi = min(R, (int)floorf((1.001f - fux) / fqx));
                const int oy_lo = max(-R, (int)ceilf((-1.001f - fuy) / fqy)), oy_hi = min(R, (int)floorf((1.001f - fuy) / fqy));
                for (int ox = ox_lo; ox <= ox_hi; ++ox) {
                    const double kx = kvac * ux + ox * qx;                                         // :395
                    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
                        const int o = p.order_map[(ox + R) * W + (oy + R)];
                        if (o < 0) continue;
                        const double ky = kvac * uy + oy * qy;                                     // :396
                        if (kx * kx + ky * ky <= kvac * kvac) {                                    // :398
                            record<STATS>(p, o, ux, uy, which, false, out.stats, out.violation);
                            if (!located) { q = locate(p, ux, uy, which); located = true; }
                            // kz (:404), phase about the cell centre (:408-409), gathers, accumulation
                            order_term<FAST>(p, o, q, Hw_x, Hw_y, kx, ky, kg, kvac, inv_kg_n, L.Z0,
                                             kx * (x - cx) + ky * (y - cy), Ex, Ey, Hx, Hy);
                        }
                    }
                }
                if (!L.plane_wave) {                                                // :453-461
                    const double path = sqrt((cx - L.source_x) * (cx - L.source_x) + (cy - L.source_y) * (cy - L.source_y) +
                                             L.source_z * L.source_z);
                    const cplx e = expi<FAST>(kvac * path);
                    Ex = Ex * e; Ey = Ey * e; Hx = Hx * e; Hy = Hy * e;
                }
            }
        }
        const size_t off = (size_t)i * out.ld + j;
        store_c(out.F[0], off, Ex, out.out_is_double);
        store_c(out.F[1], off, Ey, out.out_is_double);
        store_c(out.F[2], off, Hx, out.out_is_double);
        store_c(out.F[3], off, Hy, out.out_is_double);
    }
    // incident power through the lens (:474-477), deterministic per-block partial sums
    double v = local_power;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ double ws[NF_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NF_THREADS / 32; ++w) s += ws[w];
        out.power_block_sums[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
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

static int g_nf_minblocks = 5;   // 96 registers, 5 blocks/SM: fastest on B200 (scripts/tune_nearfield.py)
/* tuning knob: minimum resident blocks per SM the complex64 kernel is compiled for (1, 5 or 6) */
extern "C" int mlb_nearfield_tune(int min_blocks) {
    MLB_REQUIRE(min_blocks == 1 || min_blocks == 5 || min_blocks == 6, "mlb_nearfield_tune: min_blocks must be 1, 5 or 6");
    g_nf_minblocks = min_blocks;
    return MLB_OK;
}

extern "C" int mlb_nearfield_blocks(int nx, int ny) { return nx * ((ny + mlb::NF_THREADS - 1) / mlb::NF_THREADS); }

static int check_pack(const mlb_table_pack &p, const char *what) {
    MLB_REQUIRE(p.axes && p.values && p.values_f32 && p.orders, "mlb_nearfield_assemble: %s pack has NULL arrays", what);
    MLB_REQUIRE(mlb::aligned16(p.values_f32), "mlb_nearfield_assemble: %s pack values_f32 not 16-byte aligned", what);
    MLB_REQUIRE(p.order_map && p.order_radius >= 0 && p.order_radius <= 64, "mlb_nearfield_assemble: %s pack has no order map", what);
    MLB_REQUIRE(p.n_ux >= 2 && p.n_uy >= 2 && p.n_g >= 2 && p.n_orders >= 0,
                "mlb_nearfield_assemble: %s pack needs >= 2 nodes per axis (%d,%d,%d)", what, p.n_ux, p.n_uy, p.n_g);
    MLB_REQUIRE(mlb::aligned16(p.values), "mlb_nearfield_assemble: %s pack values not 16-byte aligned", what);
    return MLB_OK;
}

extern "C" int mlb_nearfield_assemble(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld,
                                      int out_is_double, double *power_block_sums, long long *stats, int want_stats,
                                      int *violation, void *stream) {
    MLB_REQUIRE(h_desc && Ex && Ey && Hx && Hy && power_block_sums && violation, "mlb_nearfield_assemble: NULL pointer");
    const mlb_lens_desc &L = *h_desc;
    MLB_REQUIRE(L.nx > 0 && L.ny > 0 && ld >= L.ny, "mlb_nearfield_assemble: bad grid (%d,%d,ld=%d)", L.nx, L.ny, ld);
    MLB_REQUIRE(L.x_pts && L.y_pts, "mlb_nearfield_assemble: NULL sample coordinates");
    MLB_REQUIRE(L.n_rings >= 1 && L.ring_boundary && L.r_center && L.grating_period && L.num_around && L.gc_index,
                "mlb_nearfield_assemble: the periphery needs >= 1 ring (nearfield.py:87-93 dereferences it)");
    MLB_REQUIRE(L.n_packs >= 1 && L.n_packs <= MLB_MAX_PACKS, "mlb_nearfield_assemble: %d collections (max %d)",
                L.n_packs, MLB_MAX_PACKS);
    for (int g = 0; g < L.n_packs; ++g)
        if (int rc = check_pack(L.packs[g], "collection")) return rc;
    if (L.n_cells > 0) {
        if (int rc = check_pack(L.hex, "hexgridset")) return rc;
        MLB_REQUIRE(L.cell_x && L.cell_y && L.cell_which && L.cell_orig && L.bin_start && L.nbx > 0 && L.nby > 0 &&
                        L.bin_size > 0,
                    "mlb_nearfield_assemble: bad centre-cell bin grid");
    }
    MLB_REQUIRE(L.plane_wave || L.source_z < 0, "mlb_nearfield_assemble: source_z must be negative (nearfield.py:84)");
    MLB_REQUIRE(L.source_pol >= 0 && L.source_pol <= 2 && !(L.plane_wave && L.source_pol == 2),
                "mlb_nearfield_assemble: bad source polarisation (nearfield.py:85, :224)");
    MLB_REQUIRE(!want_stats || stats, "mlb_nearfield_assemble: want_stats needs a stats buffer");
    MLB_REQUIRE((size_t)L.ny <= 65535u * mlb::NF_THREADS * 32u && L.nx <= 65535, "mlb_nearfield_assemble: grid too large");
    mlb::NfOut out;
    out.F[0] = Ex; out.F[1] = Ey; out.F[2] = Hx; out.F[3] = Hy;
    out.power_block_sums = power_block_sums; out.stats = stats; out.violation = violation;
    out.ld = ld; out.out_is_double = out_is_double;
    dim3 grid((L.ny + mlb::NF_THREADS - 1) / mlb::NF_THREADS, L.nx);
    const cudaStream_t st = (cudaStream_t)stream;
    if (want_stats) {
        if (out_is_double) mlb::nearfield_kernel<true, false, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, out);
        else mlb::nearfield_kernel<true, true, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, out);
    } else {
        if (out_is_double) mlb::nearfield_kernel<false, false, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, out);
        else if (g_nf_minblocks == 5) mlb::nearfield_kernel<false, true, 5><<<grid, mlb::NF_THREADS, 0, st>>>(L, out);
        else if (g_nf_minblocks == 6) mlb::nearfield_kernel<false, true, 6><<<grid, mlb::NF_THREADS, 0, st>>>(L, out);
        else mlb::nearfield_kernel<false, true, 1><<<grid, mlb::NF_THREADS, 0, st>>>(L, out);
    }
    return mlb::check_launch("mlb_nearfield_assemble");
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
