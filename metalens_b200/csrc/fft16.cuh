// Register-resident radix-16 FFT passes (included by fft.cu).
//
// The radix-4 shared-memory stages of fft.cu make one shared-memory round trip per two bits of the
// transform length; at the all-bins sizes the reference actually uses (fft2 of the whole aperture,
// nearfield_farfield.py:18-20) that is ~150 bytes of shared-memory traffic per point, more than an SM can
// move in the time HBM delivers the point.  Here every thread owns 16 points of a transform in registers
// (x[t + m N/16], m < 16), does a whole radix-16 butterfly on them, and only exchanges data through shared
// memory between butterflies: a 4096-point transform is load -> DFT16 -> exchange -> DFT16 -> exchange ->
// DFT16 -> store, i.e. two round trips instead of six.
//
// Stockham autosort formulation, stage with radix R and sub-length Ns (product of the radices before it):
//     butterfly j < N/R, k = j mod Ns:  v_r = x[j + r N/R] W_{Ns R}^{r k};  V = DFT_R(v);  y[(j-k) R + k + q Ns] = V_q
// With T = N/16 threads per transform, thread t handles the 16/R butterflies j = t + b T; their inputs are
// exactly the thread's own 16 register slots (slot b + (16/R) r), so every stage reads the same shared-memory
// positions t + m T.  All reads of a stage happen before a barrier and all writes after it, so ONE buffer
// per transform suffices.  The first stage reads global memory, the last one writes it (its outputs
// k + q Ns are consecutive in t: coalesced).  Shared-memory index n is stored at n + (n >> 4): with that
// padding the stride-16 writes of the first stage, the grouped writes of the second and the contiguous
// reads all hit 32 different banks per half-warp (no conflicts).
// Radix sequences: 256 = 16.16, 512 = 16.16.2, 1024 = 16.16.4, 2048 = 16.16.8, 4096 = 16.16.16,
// 8192 = 16.16.16.2.  Stage twiddles are powers of one table entry W_N^(k N/(Ns R)) (float64-built table),
// multiplied up along the binary expansion of r (at most 3 products deep).
#pragma once
#include <math_constants.h>

namespace mlb {

__device__ __forceinline__ int pad16(int n) { return n + (n >> 4); }

__device__ __forceinline__ float2 cmul16(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// in-register DFTs, natural order in and out
__device__ __forceinline__ void dft2r(float2 &a, float2 &b) {
    const float2 s = make_float2(a.x + b.x, a.y + b.y), d = make_float2(a.x - b.x, a.y - b.y);
    a = s; b = d;
}
__device__ __forceinline__ void dft4r(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
    const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y), d = make_float2(v1.x - v3.x, v1.y - v3.y);
    const float2 a3 = make_float2(d.y, -d.x);                                  // -i (v1 - v3)
    v0 = make_float2(a0.x + a2.x, a0.y + a2.y); v1 = make_float2(a1.x + a3.x, a1.y + a3.y);
    v2 = make_float2(a0.x - a2.x, a0.y - a2.y); v3 = make_float2(a1.x - a3.x, a1.y - a3.y);
}

template <int R>
__device__ __forceinline__ void dft_reg(float2 (&v)[R]) {
    if constexpr (R == 2) {
        dft2r(v[0], v[1]);
    } else if constexpr (R == 4) {
        dft4r(v[0], v[1], v[2], v[3]);
    } else if constexpr (R == 8) {
        // X[k] = E[k] + W8^k O[k], X[k+4] = E[k] - W8^k O[k];  E, O = DFT4 of the even / odd samples
        dft4r(v[0], v[2], v[4], v[6]);
        dft4r(v[1], v[3], v[5], v[7]);
        const float h = 0.70710678118654752440f;
        const float2 o1 = make_float2(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));     // W8^1 = (1 - i)/sqrt2
        const float2 o2 = make_float2(v[5].y, -v[5].x);                                   // W8^2 = -i
        const float2 o3 = make_float2(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));    // W8^3 = (-1 - i)/sqrt2
        const float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
        v[0] = make_float2(e0.x + o0.x, e0.y + o0.y); v[4] = make_float2(e0.x - o0.x, e0.y - o0.y);
        v[1] = make_float2(e1.x + o1.x, e1.y + o1.y); v[5] = make_float2(e1.x - o1.x, e1.y - o1.y);
        v[2] = make_float2(e2.x + o2.x, e2.y + o2.y); v[6] = make_float2(e2.x - o2.x, e2.y - o2.y);
        v[3] = make_float2(e3.x + o3.x, e3.y + o3.y); v[7] = make_float2(e3.x - o3.x, e3.y - o3.y);
    } else {
        static_assert(R == 16, "radix");
        // n = 4 n1 + n2, k = k1 + 4 k2:  X[k1 + 4 k2] = sum_n2 W4^(n2 k2) [ W16^(n2 k1) sum_n1 x[4 n1 + n2] W4^(n1 k1) ]
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) dft4r(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);   // -> A[n2][k1] in v[n2 + 4 k1]
        const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;          // cos, sin (pi/8)
        const float h = 0.70710678118654752440f;
        // W16^m = (cos(m pi/8), -sin(m pi/8));  element (n2, k1) *= W16^(n2 k1)
        v[5] = cmul16(v[5], make_float2(c1, -s1));        // n2=1,k1=1: W^1
        v[9] = cmul16(v[9], make_float2(h, -h));          // n2=1,k1=2: W^2
        v[13] = cmul16(v[13], make_float2(s1, -c1));      // n2=1,k1=3: W^3
        v[6] = cmul16(v[6], make_float2(h, -h));          // n2=2,k1=1: W^2
        v[10] = make_float2(v[10].y, -v[10].x);           // n2=2,k1=2: W^4 = -i
        v[14] = cmul16(v[14], make_float2(-h, -h));       // n2=2,k1=3: W^6
        v[7] = cmul16(v[7], make_float2(s1, -c1));        // n2=3,k1=1: W^3
        v[11] = cmul16(v[11], make_float2(-h, -h));       // n2=3,k1=2: W^6
        v[15] = cmul16(v[15], make_float2(-c1, s1));      // n2=3,k1=3: W^9 = -W^1
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4r(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // over n2 -> k2
        // v[4 k1 + k2] now holds X[k1 + 4 k2]: transpose the 4 x 4 register tile to natural order
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const float2 tmp = v[4 * p + q]; v[4 * p + q] = v[4 * q + p]; v[4 * q + p] = tmp;
            }
    }
}

// v[r] *= w^r for r = 1..R-1, powers built along the binary expansion of r
template <int R>
__device__ __forceinline__ void apply_twiddle_powers(float2 (&v)[R], float2 w1) {
    v[1] = cmul16(v[1], w1);
    if constexpr (R >= 4) {
        const float2 w2 = cmul16(w1, w1);
        v[2] = cmul16(v[2], w2);
        v[3] = cmul16(v[3], cmul16(w2, w1));
        if constexpr (R >= 8) {
            const float2 w4 = cmul16(w2, w2);
            v[4] = cmul16(v[4], w4);
            v[5] = cmul16(v[5], cmul16(w4, w1));
            const float2 w6 = cmul16(w4, w2);
            v[6] = cmul16(v[6], w6);
            v[7] = cmul16(v[7], cmul16(w6, w1));
            if constexpr (R >= 16) {
                const float2 w8 = cmul16(w4, w4);
                v[8] = cmul16(v[8], w8);
                v[9] = cmul16(v[9], cmul16(w8, w1));
                const float2 w10 = cmul16(w8, w2);
                v[10] = cmul16(v[10], w10);
                v[11] = cmul16(v[11], cmul16(w10, w1));
                const float2 w12 = cmul16(w8, w4);
                v[12] = cmul16(v[12], w12);
                v[13] = cmul16(v[13], cmul16(w12, w1));
                const float2 w14 = cmul16(w8, w6);
                v[14] = cmul16(v[14], w14);
                v[15] = cmul16(v[15], cmul16(w14, w1));
            }
        }
    }
}

// radix of stage S of a 2^LGN-point transform (see the table in the header comment) and stage count
template <int LGN, int S>
struct R16Radix {
    static constexpr int bits = LGN - 4 * S;              // bits left before stage S: radix 16 while >= 4 remain
    static constexpr int value = (bits >= 4) ? 16 : (1 << bits);
};
template <int LGN>
struct R16Stages { static constexpr int value = (LGN + 3) / 4; };

template <int R> struct Lg2 { static constexpr int value = (R == 16) ? 4 : (R == 8) ? 3 : (R == 4) ? 2 : 1; };

// the butterflies of one stage on the thread's 16 register slots; t = thread index within the transform
template <int LGN, int LGNS, int R>
__device__ __forceinline__ void r16_butterflies(float2 (&a)[16], int t, const float2 *__restrict__ tw, int tws) {
    constexpr int TR = 1 << (LGN - 4), B = 16 / R, Ns = 1 << LGNS;
#pragma unroll
    for (int b = 0; b < B; ++b) {
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = a[b + B * r];
        if constexpr (LGNS > 0) {
            const int k = (t + b * TR) & (Ns - 1);
            // W_{Ns R}^k from the table W_{N 2^tws}^t (tws > 0: table of a longer transform this one is part of)
            apply_twiddle_powers<R>(v, __ldg(tw + ((k << (LGN - LGNS - Lg2<R>::value)) << tws)));
        }
        dft_reg<R>(v);
#pragma unroll
        for (int r = 0; r < R; ++r) a[b + B * r] = v[r];
    }
}

// where output q of butterfly b lands: n = (j - k) R + k + q Ns, j = t + b TR
template <int LGN, int LGNS, int R>
__device__ __forceinline__ int r16_out_index(int t, int b, int q) {
    constexpr int TR = 1 << (LGN - 4), Ns = 1 << LGNS;
    const int j = t + b * TR, k = j & (Ns - 1);
    return ((j - k) << Lg2<R>::value) + k + (q << LGNS);
}

// runs stages S.. on the thread's registers, exchanging through `sm` (this transform's padded buffer);
// on return `a` holds the outputs of the LAST stage (slot b + B q = output q of butterfly b)
template <int LGN, int S, int LGNS, typename Sync>
__device__ __forceinline__ void r16_stages(float2 (&a)[16], int t, float2 *sm, const float2 *__restrict__ tw, int tws,
                                           Sync sync) {
    constexpr int R = R16Radix<LGN, S>::value, B = 16 / R, TR = 1 << (LGN - 4);
    r16_butterflies<LGN, LGNS, R>(a, t, tw, tws);
    if constexpr (S + 1 < R16Stages<LGN>::value) {
        if constexpr (S > 0) sync();                      // everyone has read the previous contents
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int q = 0; q < R; ++q) sm[pad16(r16_out_index<LGN, LGNS, R>(t, b, q))] = a[b + B * q];
        sync();
#pragma unroll
        for (int m = 0; m < 16; ++m) a[m] = sm[pad16(t + m * TR)];
        r16_stages<LGN, S + 1, LGNS + Lg2<R>::value>(a, t, sm, tw, tws, sync);
    }
}

// output position (within the transform) of register slot (b, q) after the last stage
template <int LGN>
__device__ __forceinline__ int r16_final_index(int t, int slot) {
    constexpr int S = R16Stages<LGN>::value - 1, R = R16Radix<LGN, S>::value, B = 16 / R;
    constexpr int LGNS = LGN - Lg2<R>::value;
    const int b = slot % B, q = slot / B;
    return r16_out_index<LGN, LGNS, R>(t, b, q);
}

template <int LGN> struct R16Threads { static constexpr int value = (LGN >= 13) ? 512 : 256; };

struct R16Args {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;                 // plain table W_N^t
    int ld_in, ld_out, n_rows, in_roll_r, in_roll_c, out_roll, s1, s2;
};

// rows: transform along the contiguous axis; L = threads / (N/16) rows per CTA.  The loader sums the s1 x s2
// aliased copies (aperture fold) and applies the input fftshift; the output fftshift is a roll at store time.
template <int LGN, int MINB = ((LGN >= 13) ? 1 : 2)>
__global__ void __launch_bounds__(R16Threads<LGN>::value, MINB) fft16_rows_kernel(const R16Args a) {
    constexpr int N = 1 << LGN, TR = N >> 4, T = R16Threads<LGN>::value, L = T / TR, PITCH = N + (N >> 4) + 1;
    extern __shared__ __align__(16) float2 fsm16[];
    const int tid = threadIdx.x, lane = tid / TR, t = tid - lane * TR;
    const int r = blockIdx.x * L + lane;
    const bool live = r < a.n_rows;
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    float2 *sm = fsm16 + lane * PITCH;
    float2 v[16];
    if (live) {
        int rs = r - a.in_roll_r; if (rs < 0) rs += a.n_rows;
        if (a.s1 == 1 && a.s2 == 1) {
            const float2 *row = in + (size_t)rs * a.ld_in;
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = __ldcs(row + ((t + m * TR - a.in_roll_c) & (N - 1)));
        } else {
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = make_float2(0.f, 0.f);
            for (int t1 = 0; t1 < a.s1; ++t1) {
                const float2 *row = in + (size_t)(rs + t1 * a.n_rows) * a.ld_in;
                for (int t2 = 0; t2 < a.s2; ++t2) {
                    const float2 *seg = row + ((size_t)t2 << LGN);
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const float2 x = __ldcs(seg + ((t + m * TR - a.in_roll_c) & (N - 1)));
                        v[m].x += x.x; v[m].y += x.y;
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = make_float2(0.f, 0.f);
    }
    r16_stages<LGN, 0, 0>(v, t, sm, a.tw, 0, [] { __syncthreads(); });
    if (live) {
        float2 *dst = out + (size_t)r * a.ld_out;
#pragma unroll
        for (int m = 0; m < 16; ++m) dst[(r16_final_index<LGN>(t, m) + a.out_roll) & (N - 1)] = v[m];
    }
}

// columns: transform along the strided axis, CL adjacent columns per CTA (8*CL-byte row segments), each column
// with its own padded shared-memory buffer.  Serves the direct column pass (g = 0) and the second pass of the
// long-transform decomposition below: sub-transform g reads rows g*in_gs + n*in_rs and stores output q at row
// (g*out_gs + q*out_rs + roll) mod Ntot.
struct R16ColArgs {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;                 // plain table W_Ntot^t of the FULL length (sub-transforms index it with a stride)
    int ld_in, ld_out, n_cols, lgNtot;
    int in_gs, in_rs, out_gs, out_rs, roll;
};

template <int LGN, int CL, int MINB = ((CL * (1 << (LGN - 4)) <= 256) ? 2 : 1)>
__global__ void __launch_bounds__(CL * (1 << (LGN - 4)), MINB) fft16_cols_kernel(const R16ColArgs a) {
    constexpr int N = 1 << LGN, TR = N >> 4, PITCH = N + (N >> 4) + 32 / CL;
    extern __shared__ __align__(16) float2 fsm16[];
    const int tid = threadIdx.x, lane = tid % CL, t = tid / CL;
    const int c = blockIdx.x * CL + lane, g = blockIdx.y;
    const bool live = c < a.n_cols;
    const float2 *__restrict__ in = pick4(a.in, blockIdx.z);
    float2 *__restrict__ out = pick4(a.out, blockIdx.z);
    float2 *sm = fsm16 + lane * PITCH;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m)
        v[m] = live ? in[(size_t)(g * a.in_gs + (t + m * TR) * a.in_rs) * a.ld_in + c] : make_float2(0.f, 0.f);
    r16_stages<LGN, 0, 0>(v, t, sm, a.tw, a.lgNtot - LGN, [] { __syncthreads(); });
    if (live) {
        const int mask = (1 << a.lgNtot) - 1;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int orow = (g * a.out_gs + r16_final_index<LGN>(t, m) * a.out_rs + a.roll) & mask;
            out[(size_t)orow * a.ld_out + c] = v[m];
        }
    }
}

// Fused column pass + radiated power on the radix-16 engine: the CTA transforms its CL columns of ALL FOUR
// fields one after the other and accumulates, per far-field point, only
//     t1 = L_phi + Z N_theta = -px Fex - py Fey + Z cy Fhx - Z cx Fhy
//     t2 = L_theta - Z N_phi = -cy Fex + cx Fey - Z px Fhx - Z py Fhy      (nearfield_farfield.py:158-167, :184)
// ((px,py) = (ux,uy)/(sin(theta)+1e-9), (cx,cy) = (px,py) uz; DC bin (:161-169): (1,0,1,0)); then
// P = k^2/(32 pi^2 Z)(|t1|^2+|t2|^2)/(uz+1e-5) * 2 (:184-189) goes straight from registers to memory, so the
// aperture sums are never stored.  The projection coefficients of the thread's 16 points are computed once
// (float64 evanescent mask / DC test as in ff_epilogue_kernel) and parked in thread-private shared memory.
struct R16PowerArgs {
    const float2 *in[4];
    const float2 *tw;
    const double *ux, *uy;
    float *P;
    double *block_sums;
    double scale;                     // pref * amp_scale^2 * 2
    float Z;
    int ld_in, ldp, n_cols, lgNtot, accumulate;
    int in_gs, in_rs, out_gs, out_rs, roll;
    int bs_stride, bs_off;            // block_sums[blockIdx.y * bs_stride + bs_off + blockIdx.x] (column strips)
    // total != NULL (single-launch shapes only): the last CTA to finish sums the n_sums block sums in the fixed order of
    // ordered_sum_256 and writes total[0] = sum * total_scale -- total_P (:74) without a second launch.  done: zeroed
    // device counter, left zero again.
    double *total;
    double total_scale;
    unsigned int *done;
    int n_sums;
};

template <int LGN, int CL>
__global__ void __launch_bounds__(CL * (1 << (LGN - 4)), 2) fft16_cols_power_kernel(const R16PowerArgs a) {
    constexpr int N = 1 << LGN, TR = N >> 4, T = CL * TR, PITCH = N + (N >> 4) + 32 / CL;
    extern __shared__ __align__(16) float2 fsm16[];
    float2 *cP = fsm16 + CL * PITCH;                               // [16][T] (px, py) of the thread's points
    float *cZ = reinterpret_cast<float *>(cP + 16 * T);            // [16][T] uz
    const int tid = threadIdx.x, lane = tid % CL, t = tid / CL;
    const int c = blockIdx.x * CL + lane, g = blockIdx.y;
    const bool live = c < a.n_cols;
    float2 *sm = fsm16 + lane * PITCH;
    const int mask = (1 << a.lgNtot) - 1;
    const double uy = live ? a.uy[c] : 0.0;
    const double uy2 = __dmul_rn(uy, uy);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int orow = (g * a.out_gs + r16_final_index<LGN>(t, m) * a.out_rs + a.roll) & mask;
        const double ux = a.ux[orow];
        // uz^2 exactly as numpy evaluates (1 - ux**2 - uy**2): bit-identical evanescent mask (:153-155)
        const double ux2 = __dmul_rn(ux, ux);
        const double uz2 = __dsub_rn(__dsub_rn(1.0, ux2), uy2);
        float px = 1.f, py = 0.f;
        if (!(ux == 0.0 && uy == 0.0)) {
            const float inv = 1.0f / ((float)sqrt(__dadd_rn(ux2, uy2)) + 1e-9f);
            px = (float)ux * inv; py = (float)uy * inv;
        }
        cP[m * T + tid] = make_float2(px, py);
        cZ[m * T + tid] = (uz2 < 0.0) ? CUDART_NAN_F : sqrtf((float)uz2);
    }
    float2 t1[16], t2[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) t1[m] = t2[m] = make_float2(0.f, 0.f);
    const float Z = a.Z;
#pragma unroll 1
    for (int f = 0; f < 4; ++f) {
        const float2 *__restrict__ in = pick4(a.in, f);
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m)
            v[m] = live ? in[(size_t)(g * a.in_gs + (t + m * TR) * a.in_rs) * a.ld_in + c] : make_float2(0.f, 0.f);
        if (f > 0) __syncthreads();                                // the previous field's exchange buffer is free
        r16_stages<LGN, 0, 0>(v, t, sm, a.tw, a.lgNtot - LGN, [] { __syncthreads(); });
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const float2 p = cP[m * T + tid];
            const float uz = cZ[m * T + tid];
            const float cx = p.x * uz, cy = p.y * uz;
            float k1, k2;
            if (f == 0) { k1 = -p.x; k2 = -cy; }                   // Ex:  L_y = -Fex
            else if (f == 1) { k1 = -p.y; k2 = cx; }               // Ey:  L_x =  Fey
            else if (f == 2) { k1 = Z * cy; k2 = -(Z * p.x); }     // Hx:  N_y =  Fhx
            else { k1 = -(Z * cx); k2 = -(Z * p.y); }              // Hy:  N_x = -Fhy
            t1[m].x = fmaf(k1, v[m].x, t1[m].x); t1[m].y = fmaf(k1, v[m].y, t1[m].y);
            t2[m].x = fmaf(k2, v[m].x, t2[m].x); t2[m].y = fmaf(k2, v[m].y, t2[m].y);
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int orow = (g * a.out_gs + r16_final_index<LGN>(t, m) * a.out_rs + a.roll) & mask;
        const float mag = t1[m].x * t1[m].x + t1[m].y * t1[m].y + t2[m].x * t2[m].x + t2[m].y * t2[m].y;
        const double p = a.scale * (double)mag / ((double)cZ[m * T + tid] + 1e-5);
        if (live) {
            float pf = (float)p;
            float *dst = a.P + (size_t)orow * a.ldp + c;
            if (a.accumulate) pf += *dst;                          // incoherent sum over sources
            *dst = pf;
            if (isfinite(pf)) sum += p;
        }
    }
    if (a.block_sums) {                                            // :74 total_P over finite bins
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __shared__ double ws[T / 32 < 8 ? 8 : T / 32];
        __shared__ int is_last;
        if ((tid & 31) == 0) ws[tid >> 5] = sum;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < T / 32; ++w) s += ws[w];
            a.block_sums[(size_t)blockIdx.y * a.bs_stride + a.bs_off + blockIdx.x] = s;
            is_last = 0;
            if (a.total) {
                __threadfence();
                is_last = (atomicAdd(a.done, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
            }
        }
        if constexpr (T == 256) {
            __syncthreads();
            if (is_last) {                                         // every block sum is visible: finish total_P here
                __threadfence();
                const double s = ordered_sum_256(a.block_sums, a.n_sums, ws);
                if (tid == 0) {
                    a.total[0] = s * a.total_scale;
                    *a.done = 0u;
                }
            }
        }
    }
}

// Long column transforms (N = 16 B, B = 256 or 512), first pass: for every n2 < B a 16-point DFT over rows
// n2 + n1 B entirely in registers, times W_N^(n2 k1), stored IN PLACE at row k1 B + n2.  No shared memory:
// a warp is 32 adjacent columns, so every access is a 256-byte row segment.  The second pass
// (fft16_cols_kernel with in_gs = B, out_rs = 16) transforms the B contiguous rows of each k1:
//     X[k1 + 16 k2] = sum_n2 W_B^(n2 k2) [ W_N^(n2 k1) sum_n1 x[B n1 + n2] W_16^(n1 k1) ]
// The host runs the two passes strip by strip (a few hundred columns of all four fields, sized to stay in L2):
// the first pass of every strip writes into the columns of the FIRST strip (dead by then), so the intermediate
// is produced and consumed in L2 and only that one strip-sized region is ever dirty.
struct R16FirstArgs {
    const float2 *in[4];
    float2 *out[4];                   // == in for an in-place pass; same row pitch
    const float2 *tw;                 // plain table W_N^t of the FULL length N
    int ld, n_cols, B;
};

__global__ void __launch_bounds__(256) fft16_cols_first_kernel(const R16FirstArgs a) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int n2 = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (c >= a.n_cols || n2 >= a.B) return;
    const size_t off = (size_t)n2 * a.ld + c;
    const float2 *__restrict__ src = pick4(a.in, blockIdx.z) + off;
    float2 *__restrict__ dst = pick4(a.out, blockIdx.z) + off;
    const size_t step = (size_t)a.B * a.ld;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = __ldcs(src + m * step);       // read once: stream through L2
    dft_reg<16>(v);
    apply_twiddle_powers<16>(v, __ldg(a.tw + n2));
#pragma unroll
    for (int m = 0; m < 16; ++m) dst[m * step] = v[m];
}

}  // namespace mlb
