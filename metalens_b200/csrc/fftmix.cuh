// Mixed-radix FFT passes, second generation (included by fft.cu): any length N = 2^a 3^b 5^c <= 8192.
//
// The reference sizes its aperture with good_fft_number() (nearfield.py:30-36: only factors 2, 3, 5), so its
// typical grid is NOT a power of two (675 = 3^3 5^2, 3375 = 3^3 5^3, ...) and fft2 of that grid
// (nearfield_farfield.py:18-20) is what a user of the reference actually runs.  Compared with the first
// mixed-radix kernels (radix 2..5 stages, integer division per point, one column per CTA for long transforms):
//   * big in-register radices -- 16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2, chosen greedily by the host -- so a
//     3375-point transform is three radix-15 stages instead of six radix-3/5 ones; the composite butterflies are
//     Cooley-Tukey products of the 2/3/4/5-point kernels with compile-time twiddles;
//   * index arithmetic by multiply-high with host-built reciprocals (no integer division in the loops);
//   * stage twiddles as powers of ONE table entry, built along the binary expansion (<= 4 products deep);
//   * long column transforms as N = A x B in two passes like the radix-16 engine: pass 1 does A-point DFTs over
//     rows n2 + B n1 entirely in registers (a warp = 32 adjacent columns, 256-byte row segments, in place),
//     pass 2 transforms the B contiguous rows of each k1 with 16-32 columns per CTA, so every global access of
//     the column pass is a full 128..256-byte segment (the old kernel read single 8-byte elements at N > 3000).
// Stockham autosort stage, radix R, sub-length Ns:  butterfly j < N/R, k = j mod Ns:
//     v_r = x[j + r N/R] W_{Ns R}^{r k};  V = DFT_R(v);  y[(j - k) R + k + q Ns] = V_q
#pragma once

namespace mlb {

// W_R^m = (mixw_re<R>(m), mixw_im<R>(m)) for the composite radices; m is a compile-time constant after unrolling
template <int R>
__host__ __device__ constexpr float mixw_re(int m) {
    if constexpr (R == 6) { constexpr float t[6] = {1.0f, 0.5f, -0.5f, -1.0f, -0.5f, 0.5f}; return t[m]; }
    if constexpr (R == 9) { constexpr float t[9] = {1.0f, 0.766044443f, 0.173648178f, -0.5f, -0.939692621f, -0.939692621f, -0.5f, 0.173648178f, 0.766044443f}; return t[m]; }
    if constexpr (R == 10) { constexpr float t[10] = {1.0f, 0.809016994f, 0.309016994f, -0.309016994f, -0.809016994f, -1.0f, -0.809016994f, -0.309016994f, 0.309016994f, 0.809016994f}; return t[m]; }
    if constexpr (R == 12) { constexpr float t[12] = {1.0f, 0.866025404f, 0.5f, 6.123234e-17f, -0.5f, -0.866025404f, -1.0f, -0.866025404f, -0.5f, -1.8369702e-16f, 0.5f, 0.866025404f}; return t[m]; }
    if constexpr (R == 15) { constexpr float t[15] = {1.0f, 0.913545458f, 0.669130606f, 0.309016994f, -0.104528463f, -0.5f, -0.809016994f, -0.978147601f, -0.978147601f, -0.809016994f, -0.5f, -0.104528463f, 0.309016994f, 0.669130606f, 0.913545458f}; return t[m]; }
    return 1.f;
}
template <int R>
__host__ __device__ constexpr float mixw_im(int m) {
    if constexpr (R == 6) { constexpr float t[6] = {-0.0f, -0.866025404f, -0.866025404f, -1.2246468e-16f, 0.866025404f, 0.866025404f}; return t[m]; }
    if constexpr (R == 9) { constexpr float t[9] = {-0.0f, -0.64278761f, -0.984807753f, -0.866025404f, -0.342020143f, 0.342020143f, 0.866025404f, 0.984807753f, 0.64278761f}; return t[m]; }
    if constexpr (R == 10) { constexpr float t[10] = {-0.0f, -0.587785252f, -0.951056516f, -0.951056516f, -0.587785252f, -1.2246468e-16f, 0.587785252f, 0.951056516f, 0.951056516f, 0.587785252f}; return t[m]; }
    if constexpr (R == 12) { constexpr float t[12] = {-0.0f, -0.5f, -0.866025404f, -1.0f, -0.866025404f, -0.5f, -1.2246468e-16f, 0.5f, 0.866025404f, 1.0f, 0.866025404f, 0.5f}; return t[m]; }
    if constexpr (R == 15) { constexpr float t[15] = {-0.0f, -0.406736643f, -0.743144825f, -0.951056516f, -0.994521895f, -0.866025404f, -0.587785252f, -0.207911691f, 0.207911691f, 0.587785252f, 0.866025404f, 0.994521895f, 0.951056516f, 0.743144825f, 0.406736643f}; return t[m]; }
    return 0.f;
}

template <int R> struct MixSplit { static constexpr int A = R, B = 1; };      // base radices: no split
template <> struct MixSplit<6> { static constexpr int A = 3, B = 2; };
template <> struct MixSplit<9> { static constexpr int A = 3, B = 3; };
template <> struct MixSplit<10> { static constexpr int A = 5, B = 2; };
template <> struct MixSplit<12> { static constexpr int A = 4, B = 3; };
template <> struct MixSplit<15> { static constexpr int A = 5, B = 3; };

template <int R>
__device__ __forceinline__ void mix_base(float2 (&t)[R]) {
    if constexpr (R == 8 || R == 16) dft_reg<R>(t);
    else dft_small<R>(t);
}

// register slot that holds output q of dft_mix<R>
template <int R>
__device__ __forceinline__ constexpr int mix_slot(int q) {
    return MixSplit<R>::B == 1 ? q : (q / MixSplit<R>::A) + MixSplit<R>::B * (q % MixSplit<R>::A);
}

// in-register R-point DFT; output q is left in slot mix_slot<R>(q).  Composite R = A B (n = B n1 + n2, k = k1 + A k2):
//     X[k1 + A k2] = sum_n2 W_B^(n2 k2) [ W_R^(n2 k1) sum_n1 x[B n1 + n2] W_A^(n1 k1) ]
template <int R>
__device__ __forceinline__ void dft_mix(float2 (&v)[R]) {
    constexpr int A = MixSplit<R>::A, B = MixSplit<R>::B;
    if constexpr (B == 1) {
        mix_base<R>(v);
    } else {
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) {                      // A-point DFTs over n1 -> k1, slot n2 + B k1
            float2 t[A];
#pragma unroll
            for (int n1 = 0; n1 < A; ++n1) t[n1] = v[n2 + B * n1];
            mix_base<A>(t);
#pragma unroll
            for (int k1 = 0; k1 < A; ++k1) {
                const int m = (n2 * k1) % R;
                v[n2 + B * k1] = (m == 0) ? t[k1] : cmul16(t[k1], make_float2(mixw_re<R>(m), mixw_im<R>(m)));
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {                      // B-point DFTs over n2 -> k2, slot k2 + B k1
            float2 t[B];
#pragma unroll
            for (int n2 = 0; n2 < B; ++n2) t[n2] = v[n2 + B * k1];
            mix_base<B>(t);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) v[k2 + B * k1] = t[k2];
        }
    }
}

// v[r] *= w^r for r = 1..R-1.  Powers two-level, w^r = (w^4)^(r/4) w^(r%4): six live complex values whatever R is
// (a full table of R powers costs 2R registers at R = 16), every power at most 4 products deep.
template <int R>
__device__ __forceinline__ void mix_twiddle(float2 (&v)[R], float2 w1) {
    if constexpr (R == 2) {
        v[1] = cmul16(v[1], w1);
    } else {
        const float2 w2 = cmul16(w1, w1), w3 = cmul16(w2, w1);
        v[1] = cmul16(v[1], w1);
        v[2] = cmul16(v[2], w2);
        if constexpr (R > 3) v[3] = cmul16(v[3], w3);
        if constexpr (R > 4) {
            float2 hi = cmul16(w2, w2);                        // w^4, then w^8, w^12
            const float2 w4 = hi;
#pragma unroll
            for (int a4 = 4; a4 < R; a4 += 4) {
                v[a4] = cmul16(v[a4], hi);
                if (a4 + 1 < R) v[a4 + 1] = cmul16(v[a4 + 1], cmul16(hi, w1));
                if (a4 + 2 < R) v[a4 + 2] = cmul16(v[a4 + 2], cmul16(hi, w2));
                if (a4 + 3 < R) v[a4 + 3] = cmul16(v[a4 + 3], cmul16(hi, w3));
                if (a4 + 4 < R) hi = cmul16(hi, w4);
            }
        }
    }
}
// same for outputs held in the slot order of dft_mix<A>: output k1 (in slot mix_slot<A>(k1)) *= w^k1
template <int A>
__device__ __forceinline__ void mix_twiddle_slots(float2 (&v)[A], float2 w1) {
    float2 t[A];
#pragma unroll
    for (int k = 0; k < A; ++k) t[k] = v[mix_slot<A>(k)];
    mix_twiddle<A>(t, w1);
#pragma unroll
    for (int k = 0; k < A; ++k) v[mix_slot<A>(k)] = t[k];
}

constexpr int MIX2_MAX_STAGES = 8;
// shared-memory position of element n of a transform.  The first stage writes with stride R: an odd R spreads
// over the banks by itself (and a pad every 16 elements would turn stride 15 into stride ~16: every thread on one
// bank), an even R needs one pad element every 16 (stride 16 -> 17).  The host picks sh = 4 (pad) when the first
// radix is even, sh = 30 (n >> 30 = 0: no pad) when it is odd.
__host__ __device__ __forceinline__ constexpr int mixpad(int n, int sh) { return n + (n >> sh); }

struct Mix2Args {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;                 // plain table W_Ntot^t, t < Ntot
    int ld_in, ld_out, N, other, lanes, lg_lanes;
    int in_roll_r, in_roll_c, out_roll, s1, s2;          // rows: fold factors and fftshift rolls
    int tw_mul;                                          // Ntot / N (a sub-transform indexes the long table with a stride)
    int in_gs, in_rs, out_gs, out_rs, roll, Ntot;        // columns: in row = g in_gs + n in_rs, out row = (g out_gs + q out_rs + roll) mod Ntot
    int nstage, pad_sh;
    int radix[MIX2_MAX_STAGES];
    unsigned magic_ns[MIX2_MAX_STAGES];                  // ceil(2^32 / Ns) of the stage (unused while Ns == 1)
    unsigned magic_per[MIX2_MAX_STAGES];                 // ceil(2^32 / (N / R)) of the stage (rows: lane of a flat index)
};

// one stage on `lanes` transforms in shared memory.  COLS: element (lane, n) at mixpad(n, a.pad_sh)*lanes + lane, flat index =
// j*lanes + lane (lanes a power of two); rows: element at lane*pitch + mixpad(n, a.pad_sh), flat index = lane*per + j.
template <int R, bool COLS>
__device__ __forceinline__ void mix2_stage(const float2 *__restrict__ x, float2 *__restrict__ y, const Mix2Args &a, int s,
                                           int Ns, int pitch) {
    const int N = a.N, per = N / R, total = per * a.lanes;
    const int twstep = (N / (Ns * R)) * a.tw_mul;
    const unsigned mns = a.magic_ns[s], mper = a.magic_per[s];
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int lane, j;
        if (COLS) { lane = idx & (a.lanes - 1); j = idx >> a.lg_lanes; }
        else { lane = (per == 1) ? idx : (int)__umulhi((unsigned)idx, mper); j = idx - lane * per; }
        const int k = (Ns == 1) ? 0 : j - (int)__umulhi((unsigned)j, mns) * Ns;
        const int base = COLS ? lane : lane * pitch, mul = COLS ? a.lanes : 1;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = x[base + mixpad(j + r * per, a.pad_sh) * mul];
        if (Ns > 1) mix_twiddle<R>(v, __ldg(a.tw + k * twstep));
        dft_mix<R>(v);
        const int o0 = (j - k) * R + k;
#pragma unroll
        for (int q = 0; q < R; ++q) y[base + mixpad(o0 + q * Ns, a.pad_sh) * mul] = v[mix_slot<R>(q)];
    }
}

// all stages, ping-pong between the two buffers; returns which buffer holds the result
template <bool COLS>
__device__ __forceinline__ int mix2_fft(float2 *buf0, float2 *buf1, const Mix2Args &a, int pitch) {
    int cur = 0, Ns = 1;
    for (int s = 0; s < a.nstage; ++s) {
        __syncthreads();
        const float2 *x = cur ? buf1 : buf0;
        float2 *y = cur ? buf0 : buf1;
        switch (a.radix[s]) {
            case 16: mix2_stage<16, COLS>(x, y, a, s, Ns, pitch); break;
            case 15: mix2_stage<15, COLS>(x, y, a, s, Ns, pitch); break;
            case 12: mix2_stage<12, COLS>(x, y, a, s, Ns, pitch); break;
            case 10: mix2_stage<10, COLS>(x, y, a, s, Ns, pitch); break;
            case 9: mix2_stage<9, COLS>(x, y, a, s, Ns, pitch); break;
            case 8: mix2_stage<8, COLS>(x, y, a, s, Ns, pitch); break;
            case 6: mix2_stage<6, COLS>(x, y, a, s, Ns, pitch); break;
            case 5: mix2_stage<5, COLS>(x, y, a, s, Ns, pitch); break;
            case 4: mix2_stage<4, COLS>(x, y, a, s, Ns, pitch); break;
            case 3: mix2_stage<3, COLS>(x, y, a, s, Ns, pitch); break;
            default: mix2_stage<2, COLS>(x, y, a, s, Ns, pitch); break;
        }
        Ns *= a.radix[s];
        cur ^= 1;
    }
    __syncthreads();
    return cur;
}

// rows: `lanes` rows per CTA; the loader folds the s1 x s2 aliased copies and applies the input fftshift
__global__ void __launch_bounds__(256) mix2_rows_kernel(const Mix2Args a) {
    extern __shared__ __align__(16) float2 fsm_mix[];
    const int N = a.N, L = a.lanes, P = mixpad(N, a.pad_sh) + 1;
    float2 *buf0 = fsm_mix, *buf1 = fsm_mix + (size_t)L * P;
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    const int row0 = blockIdx.x * L;
    for (int lane = 0; lane < L; ++lane) {
        const int r = row0 + lane;
        if (r >= a.other) {
            for (int n = threadIdx.x; n < N; n += blockDim.x) buf0[lane * P + mixpad(n, a.pad_sh)] = make_float2(0.f, 0.f);
            continue;
        }
        int rs = r - a.in_roll_r; if (rs < 0) rs += a.other;
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            int cs = n - a.in_roll_c; if (cs < 0) cs += N;
            float re = 0.f, im = 0.f;
            for (int t1 = 0; t1 < a.s1; ++t1) {
                const float2 *row = in + (size_t)(rs + t1 * a.other) * a.ld_in + cs;
                for (int t2 = 0; t2 < a.s2; ++t2) {
                    const float2 v = __ldcs(row + (size_t)t2 * N);
                    re += v.x; im += v.y;
                }
            }
            buf0[lane * P + mixpad(n, a.pad_sh)] = make_float2(re, im);
        }
    }
    const int cur = mix2_fft<false>(buf0, buf1, a, P);
    const float2 *res = cur ? buf1 : buf0;
    for (int lane = 0; lane < L; ++lane) {
        const int r = row0 + lane;
        if (r >= a.other) break;
        float2 *dst = out + (size_t)r * a.ld_out;
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            int q = n - a.out_roll; if (q < 0) q += N;
            dst[n] = res[lane * P + mixpad(q, a.pad_sh)];
        }
    }
}

// columns: `lanes` (power of two) adjacent columns per CTA, shared-memory element (lane, n) at mixpad(n, a.pad_sh)*lanes + lane, so
// global row segments and shared-memory accesses are both contiguous in the lane.  blockIdx.y = sub-transform g
// of the two-pass decomposition (0 for a direct pass), blockIdx.z = field.
__global__ void __launch_bounds__(256) mix2_cols_kernel(const Mix2Args a) {
    extern __shared__ __align__(16) float2 fsm_mix[];
    const int N = a.N, L = a.lanes;
    float2 *buf0 = fsm_mix, *buf1 = fsm_mix + (size_t)L * (mixpad(N, a.pad_sh) + 1);
    const float2 *__restrict__ in = pick4(a.in, blockIdx.z);
    float2 *__restrict__ out = pick4(a.out, blockIdx.z);
    const int c0 = blockIdx.x * L, g = blockIdx.y, total = L * N;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int lane = idx & (L - 1), n = idx >> a.lg_lanes;
        const int c = c0 + lane;
        buf0[mixpad(n, a.pad_sh) * L + lane] = (c < a.other) ? in[(size_t)(g * a.in_gs + n * a.in_rs) * a.ld_in + c] : make_float2(0.f, 0.f);
    }
    const int cur = mix2_fft<true>(buf0, buf1, a, 0);
    const float2 *res = cur ? buf1 : buf0;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int lane = idx & (L - 1), q = idx >> a.lg_lanes;
        const int c = c0 + lane;
        if (c < a.other) {
            int orow = g * a.out_gs + q * a.out_rs + a.roll;
            if (orow >= a.Ntot) orow -= a.Ntot;
            out[(size_t)orow * a.ld_out + c] = res[mixpad(q, a.pad_sh) * L + lane];
        }
    }
}

// first pass of a long column transform N = A B: for every n2 < B an A-point DFT over rows n2 + n1 B in registers,
// times W_N^(n2 k1), stored at row k1 B + n2 (in place when out == in); no shared memory
struct Mix2FirstArgs {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;                 // plain table W_N^t of the FULL length
    int ld, n_cols, B;
};

template <int A>
__global__ void __launch_bounds__(256) mix2_cols_first_kernel(const Mix2FirstArgs a) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int n2 = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (c >= a.n_cols || n2 >= a.B) return;
    const size_t off = (size_t)n2 * a.ld + c;
    const float2 *__restrict__ src = pick4(a.in, blockIdx.z) + off;
    float2 *__restrict__ dst = pick4(a.out, blockIdx.z) + off;
    const size_t step = (size_t)a.B * a.ld;
    float2 v[A];
#pragma unroll
    for (int m = 0; m < A; ++m) v[m] = __ldcs(src + m * step);
    dft_mix<A>(v);
    mix_twiddle_slots<A>(v, __ldg(a.tw + n2));
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) dst[k1 * step] = v[mix_slot<A>(k1)];
}

// ---- register-resident variant: ONE butterfly per thread and stage ------------------------------------------
// When lanes * N / R <= blockDim for every stage, a thread keeps its butterfly in registers across the barrier,
// so the first stage reads global memory directly, the last one writes it directly, and the stages in between
// exchange through ONE shared-memory buffer (read all -> barrier -> write all), as in the radix-16 engine:
// half the shared memory (7 resident CTAs per SM at 3375 points instead of 3) and two shared-memory passes fewer.
template <int R, bool COLS>
__device__ __forceinline__ void mix2_reg_stage(const Mix2Args &a, int s, int Ns, float2 *__restrict__ sm, int pitch,
                                               const float2 *__restrict__ in, float2 *__restrict__ out, int first_unit,
                                               int g) {
    const int N = a.N, per = N / R, total = per * a.lanes;
    const bool first = (s == 0), last = (s == a.nstage - 1);
    const int idx = threadIdx.x;
    const bool act = idx < total;
    int lane = 0, j = 0;
    if (COLS) { lane = idx & (a.lanes - 1); j = idx >> a.lg_lanes; }
    else { lane = (per == 1) ? idx : (int)__umulhi((unsigned)idx, a.magic_per[s]); j = idx - lane * per; }
    const int unit = first_unit + lane;                   // row (rows kernel) or column (columns kernel) of this lane
    const bool live = act && unit < a.other;
    const int base = COLS ? lane : lane * pitch, mul = COLS ? a.lanes : 1;
    float2 v[R];
    if (first) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = make_float2(0.f, 0.f);
        if (live) {
            if (COLS) {
                // element offsets fit 32 bits (N * pitch < 2^27): one base pointer, short-lived addresses
                const float2 *src = in + ((g * a.in_gs + j * a.in_rs) * a.ld_in + unit);
                const int stride = per * a.in_rs * a.ld_in;
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = src[r * stride];
            } else {
                int rs = unit - a.in_roll_r; if (rs < 0) rs += a.other;
                if (a.s1 == 1 && a.s2 == 1) {
                    const float2 *row = in + (size_t)rs * a.ld_in;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int cs = j + r * per - a.in_roll_c; if (cs < 0) cs += N;
                        v[r] = __ldcs(row + cs);
                    }
                } else {
#pragma unroll 1
                    for (int t1 = 0; t1 < a.s1; ++t1) {
                        const float2 *row = in + (size_t)(rs + t1 * a.other) * a.ld_in;
#pragma unroll 1
                        for (int t2 = 0; t2 < a.s2; ++t2) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                int cs = j + r * per - a.in_roll_c; if (cs < 0) cs += N;
                                const float2 x = __ldcs(row + (size_t)t2 * N + cs);
                                v[r].x += x.x; v[r].y += x.y;
                            }
                        }
                    }
                }
            }
        }
    } else {
        if (act) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = sm[base + mixpad(j + r * per, a.pad_sh) * mul];
        }
        __syncthreads();                                  // everyone has read the previous stage's output
    }
    const int k = (Ns == 1) ? 0 : j - (int)__umulhi((unsigned)j, a.magic_ns[s]) * Ns;
    if (act) {
        if (Ns > 1) mix_twiddle<R>(v, __ldg(a.tw + k * ((N / (Ns * R)) * a.tw_mul)));
        dft_mix<R>(v);
        const int o0 = (j - k) * R + k;
        if (!last) {
#pragma unroll
            for (int q = 0; q < R; ++q) sm[base + mixpad(o0 + q * Ns, a.pad_sh) * mul] = v[mix_slot<R>(q)];
        } else if (live) {
            if (COLS) {
                float2 *dst = out + unit;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    int orow = g * a.out_gs + (o0 + q * Ns) * a.out_rs + a.roll;
                    if (orow >= a.Ntot) orow -= a.Ntot;
                    dst[orow * a.ld_out] = v[mix_slot<R>(q)];
                }
            } else {
                float2 *dst = out + (size_t)unit * a.ld_out;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    int n = o0 + q * Ns + a.out_roll; if (n >= N) n -= N;
                    dst[n] = v[mix_slot<R>(q)];
                }
            }
        }
    }
    if (!last) __syncthreads();
}

template <bool COLS, int MINB>
__global__ void __launch_bounds__(256, MINB) mix2_reg_kernel(const Mix2Args a) {
    extern __shared__ __align__(16) float2 fsm_mix[];
    const float2 *__restrict__ in = pick4(a.in, COLS ? blockIdx.z : blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, COLS ? blockIdx.z : blockIdx.y);
    const int first_unit = blockIdx.x * a.lanes, g = COLS ? blockIdx.y : 0, pitch = mixpad(a.N, a.pad_sh) + 1;
    int Ns = 1;
    for (int s = 0; s < a.nstage; ++s) {
        switch (a.radix[s]) {
            case 16: mix2_reg_stage<16, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 15: mix2_reg_stage<15, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 12: mix2_reg_stage<12, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 10: mix2_reg_stage<10, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 9: mix2_reg_stage<9, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 8: mix2_reg_stage<8, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 6: mix2_reg_stage<6, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 5: mix2_reg_stage<5, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 4: mix2_reg_stage<4, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            case 3: mix2_reg_stage<3, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
            default: mix2_reg_stage<2, COLS>(a, s, Ns, fsm_mix, pitch, in, out, first_unit, g); break;
        }
        Ns *= a.radix[s];
    }
}

// ---- compile-time plans of the register variant ----------------------------------------------------------------------
// The register kernels serve few plans -- rows: the 11 three-stage lengths 2160..3840 (one row per CTA); second column
// pass: 14 two-stage sub-lengths 30..240 with 16 columns per CTA -- so every plan gets its own instantiation: radices,
// sub-lengths, lane count and padding are constants, the per-stage switch, the multiply-high index arithmetic and the
// spills of the generic kernel (128 registers and still 2 KB of spill stores for rows) are gone, and a stage is the
// same code as mix2_reg_stage with those constants folded in.
template <int L> struct MixLg { static constexpr int value = (L <= 1) ? 0 : 1 + MixLg<L / 2>::value; };

template <int R, bool COLS, int N, int NS, int LANES, int SH, bool FIRST, bool LAST>
__device__ __forceinline__ void mix2_ct_stage(const Mix2Args &a, float2 *__restrict__ sm, const float2 *__restrict__ in,
                                              float2 *__restrict__ out, int first_unit, int g) {
    constexpr int per = N / R, total = per * LANES, pitch = mixpad(N, SH) + 1, mul = COLS ? LANES : 1;
    static_assert(total <= 256 && N % (NS * R) == 0, "one butterfly per thread and stage");
    const int idx = threadIdx.x;
    const bool act = idx < total;
    int lane, j;
    if constexpr (COLS) { lane = idx & (LANES - 1); j = idx >> MixLg<LANES>::value; }
    else { lane = idx / per; j = idx - lane * per; }
    const int unit = first_unit + lane;                   // row (rows kernel) or column (columns kernel) of this lane
    const bool live = act && unit < a.other;
    const int base = COLS ? lane : lane * pitch;
    float2 v[R];
    if constexpr (FIRST) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = make_float2(0.f, 0.f);
        if (live) {
            if constexpr (COLS) {
                const float2 *src = in + ((g * a.in_gs + j * a.in_rs) * a.ld_in + unit);
                const int stride = per * a.in_rs * a.ld_in;
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = src[r * stride];
            } else {
                int rs = unit - a.in_roll_r; if (rs < 0) rs += a.other;
                if (a.s1 == 1 && a.s2 == 1) {
                    const float2 *row = in + (size_t)rs * a.ld_in;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int cs = j + r * per - a.in_roll_c; if (cs < 0) cs += N;
                        v[r] = __ldcs(row + cs);
                    }
                } else {
#pragma unroll 1
                    for (int t1 = 0; t1 < a.s1; ++t1) {
                        const float2 *row = in + (size_t)(rs + t1 * a.other) * a.ld_in;
#pragma unroll 1
                        for (int t2 = 0; t2 < a.s2; ++t2) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                int cs = j + r * per - a.in_roll_c; if (cs < 0) cs += N;
                                const float2 x = __ldcs(row + (size_t)t2 * N + cs);
                                v[r].x += x.x; v[r].y += x.y;
                            }
                        }
                    }
                }
            }
        }
    } else {
        if (act) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = sm[base + mixpad(j + r * per, SH) * mul];
        }
        __syncthreads();                                  // everyone has read the previous stage's output
    }
    const int k = (NS == 1) ? 0 : j % NS;
    if (act) {
        if constexpr (NS > 1) mix_twiddle<R>(v, __ldg(a.tw + k * ((N / (NS * R)) * a.tw_mul)));
        dft_mix<R>(v);
        const int o0 = (j - k) * R + k;
        if constexpr (!LAST) {
#pragma unroll
            for (int q = 0; q < R; ++q) sm[base + mixpad(o0 + q * NS, SH) * mul] = v[mix_slot<R>(q)];
        } else if (live) {
            if constexpr (COLS) {
                float2 *dst = out + unit;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    int orow = g * a.out_gs + (o0 + q * NS) * a.out_rs + a.roll;
                    if (orow >= a.Ntot) orow -= a.Ntot;
                    dst[orow * a.ld_out] = v[mix_slot<R>(q)];
                }
            } else {
                float2 *dst = out + (size_t)unit * a.ld_out;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    int n = o0 + q * NS + a.out_roll; if (n >= N) n -= N;
                    dst[n] = v[mix_slot<R>(q)];
                }
            }
        }
    }
    if constexpr (!LAST) __syncthreads();
}

constexpr int MIX2_CT_COL_LANES = 16;
// N = R0 R1 R2 (R2 = 1: two stages); rows: one row per CTA, columns: 16 adjacent columns per CTA; a.lanes, a.pad_sh and
// a.radix[] must describe exactly this plan (the host looks the instantiation up by them)
template <bool COLS, int MINB, int R0, int R1, int R2>
__global__ void __launch_bounds__(256, MINB) mix2_ct_kernel(const Mix2Args a) {
    constexpr int N = R0 * R1 * R2, LANES = COLS ? MIX2_CT_COL_LANES : 1, SH = (R0 & 1) ? 30 : 4;
    extern __shared__ __align__(16) float2 fsm_mix[];
    const float2 *__restrict__ in = pick4(a.in, COLS ? blockIdx.z : blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, COLS ? blockIdx.z : blockIdx.y);
    const int first_unit = blockIdx.x * LANES, g = COLS ? blockIdx.y : 0;
    mix2_ct_stage<R0, COLS, N, 1, LANES, SH, true, false>(a, fsm_mix, in, out, first_unit, g);
    mix2_ct_stage<R1, COLS, N, R0, LANES, SH, false, R2 == 1>(a, fsm_mix, in, out, first_unit, g);
    if constexpr (R2 > 1) mix2_ct_stage<R2, COLS, N, R0 * R1, LANES, SH, false, true>(a, fsm_mix, in, out, first_unit, g);
}

}  // namespace mlb
