// FFT passes for FFT-bin far-field grids: this file holds the radix-4 shared-memory kernels, the
// TMA-fed row pass and the host dispatch; fft16.cuh the register-resident radix-16 kernels (1024..8192 points),
// fftmix.cuh the big-radix engine for the non-power-of-two good_fft_number() lengths.
//
// The reference's own algorithm is fft2(fftshift(J)) (nearfield_farfield.py:18-20).  For grids that
// are (a stride of) the FFT bins, the aperture sum is a 2-D DFT of the (folded) aperture, done here
// as two streaming passes -- rows, then columns -- each a Stockham autosort radix-4 (+ one radix-2)
// FFT entirely inside shared memory: one read and one write of the data per pass.
//   * the row pass folds the aperture while loading (sum of the s1*s2 aliased samples per point,
//     see fold.cu), so for a strided grid the full aperture is read from HBM exactly once and the
//     folded aperture never goes to memory: this kernel is the HBM-bound hot kernel of NF->FF;
//   * for 256..2048-point rows that pass is a persistent TMA-fed producer/consumer pipeline
//     (fft_rows_tma_kernel: 94 % of the measured HBM copy bandwidth on B200); other sizes use the
//     thread-issued-load kernels (fft_rows_kernel), compile-time sized for 256..8192 points;
//   * fftshift of input and output is index arithmetic ("rolls") at load/store time;
//   * twiddles come from tables built with float64 phases, re-ordered per Stockham stage so that the
//     shared-memory reads are conflict-free.
// The kernels of this file serve powers of two (index math is shifts and masks); the first-generation radix 2..5
// mixed kernels and the counter-driven row distribution of round 1 were measured slower and have been removed.
#include <math_constants.h>

#include <string>

#include "common.cuh"
#include "fft16.cuh"

namespace mlb {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// One Stockham stage of radix R over `lanes` independent transforms of length 2^lgN held in shared
// memory; element n of transform `lane` lives at lane*pitch + n (pitch >= N, even).
template <int R>
__device__ __forceinline__ void stockham_stage(const float2 *__restrict__ x, float2 *__restrict__ y, int lgN, int lgNs,
                                               int lanes, int pitch, const float2 *__restrict__ tw) {
    constexpr int lgR = (R == 4) ? 2 : 1;
    const int lgPer = lgN - lgR;
    const int per = 1 << lgPer;
    const int total = lanes << lgPer;
    const int Ns = 1 << lgNs;
    const int lgTstep = lgN - lgNs - lgR;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int lane = idx >> lgPer, j = idx & (per - 1);
        const int k = j & (Ns - 1);
        const int base_out = ((j - k) << lgR) + k;             // expand(j, Ns, R)
        const float2 *xl = x + lane * pitch;
        float2 *yl = y + lane * pitch;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = xl[j + r * per];
            if (r > 0 && lgNs > 0) v[r] = cmulf(v[r], tw[(r * k) << lgTstep]);
        }
        if (R == 4) {
            const float2 a0 = caddf(v[0], v[2]), a1 = csubf(v[0], v[2]);
            const float2 a2 = caddf(v[1], v[3]), a3 = mul_mi(csubf(v[1], v[3]));
            v[0] = caddf(a0, a2); v[1] = caddf(a1, a3); v[2] = csubf(a0, a2); v[3] = csubf(a1, a3);
        } else {
            const float2 a0 = caddf(v[0], v[1]), a1 = csubf(v[0], v[1]);
            v[0] = a0; v[1] = a1;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) yl[base_out + (r << lgNs)] = v[r];
    }
}

// Runs the stages from sub-transform length 2^lgNs0 on (data in buf0); returns the buffer (0/1) with
// the result.  The first stage (lgNs = 0, all twiddles 1) is done by the loaders in registers.
__device__ __forceinline__ int fft_in_smem(float2 *buf0, float2 *buf1, int lgN, int lgNs0, int lanes, int pitch,
                                           const float2 *tw) {
    int cur = 0, lgNs = lgNs0;
    while (lgNs + 2 <= lgN) {
        __syncthreads();
        if (cur == 0) stockham_stage<4>(buf0, buf1, lgN, lgNs, lanes, pitch, tw);
        else stockham_stage<4>(buf1, buf0, lgN, lgNs, lanes, pitch, tw);
        cur ^= 1;
        lgNs += 2;
    }
    if (lgNs < lgN) {
        __syncthreads();
        if (cur == 0) stockham_stage<2>(buf0, buf1, lgN, lgNs, lanes, pitch, tw);
        else stockham_stage<2>(buf1, buf0, lgN, lgNs, lanes, pitch, tw);
        cur ^= 1;
    }
    __syncthreads();
    return cur;
}

// first-stage butterfly on 4 (or 2) loaded values, twiddle-free
__device__ __forceinline__ void bfly4(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
    const float2 a0 = caddf(v0, v2), a1 = csubf(v0, v2);
    const float2 a2 = caddf(v1, v3), a3 = mul_mi(csubf(v1, v3));
    v0 = caddf(a0, a2); v1 = caddf(a1, a3); v2 = csubf(a0, a2); v3 = csubf(a1, a3);
}

struct FftArgs {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;
    int ld_in, ld_out, lgN, other, lanes, in_roll_r, in_roll_c, out_roll, s1, s2, plain_loader, transpose_out, evict_first;
};

// Distributed 2-D transform (mlb_fft_rows_scatter): output row r of this rank is row (out_row0 + r) mod rows_total of
// the intermediate, and its column slab p goes to rank p's peer-mapped buffer out[p][field].  world == 0: off.
struct RowScatter {
    float2 *out[MLB_MAX_PEERS][4];
    int world, lg_slab, out_row0, rows_total;
    int lg_block, row_stride;         // local row r -> row out_row0 + (r >> lg_block) * row_stride + (r & (block - 1))
};

// folded, fftshift-rolled input sample(s) of row r at position n (VEC consecutive positions).
// Kept deliberately light (4 loads in flight, ~32 registers/thread): measured on B200, a deeper
// per-thread load queue costs more in occupancy than it gains (scripts/tune_fft_rows.py).
template <int VEC>
__device__ __forceinline__ void load_folded(const FftArgs &a, const float2 *__restrict__ in, int rs, int n, int N,
                                            float (&acc)[2 * VEC]) {
#pragma unroll
    for (int v = 0; v < 2 * VEC; ++v) acc[v] = 0.f;
    int cs = n - a.in_roll_c; if (cs < 0) cs += N;
    for (int t1 = 0; t1 < a.s1; ++t1) {
        const float2 *row = in + (size_t)(rs + t1 * a.other) * a.ld_in + cs;
#pragma unroll 4
        for (int t2 = 0; t2 < a.s2; ++t2) {
            if (VEC == 2) {
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(row + ((size_t)t2 << a.lgN)));
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            } else {
                const float2 v = __ldcs(row + ((size_t)t2 << a.lgN));
                acc[0] += v.x; acc[1] += v.y;
            }
        }
    }
}

// Threads per CTA of the compile-time-sized kernels: transforms of >= 4096 points leave room for only
// one CTA per SM (shared memory), so those CTAs are 1024 threads wide to keep enough loads in flight.
template <int LGN> struct FftThreads { static constexpr int value = (LGN >= 12) ? 1024 : 256; };

// ---- compile-time-sized stages: all index math folds into constants, loops fully unrolled ----------
template <int R, int LGN, int LGNS, int LANES, int PITCH, int T = FftThreads<LGN>::value>
__device__ __forceinline__ void stage_ct(const float2 *__restrict__ x, float2 *__restrict__ y,
                                         const float2 *__restrict__ tw) {
    constexpr int lgR = (R == 4) ? 2 : 1;
    constexpr int lgPer = LGN - lgR, per = 1 << lgPer, total = per * LANES, Ns = 1 << LGNS;
    // staged twiddle table (built by mlb_fft_twiddle after the plain one): for the radix-4 stage with
    // sub-length Ns = 4^s the entries W^(r k), r = 1..3, k < Ns sit at [4^s - 1 + 3k + (r-1)]; for a final
    // radix-2 stage at [Ns - 1 + k].  Consecutive threads read consecutive 24-byte triples instead of
    // one strided word of the length-N table: no shared-memory bank conflicts on the twiddle loads.
    constexpr int tbase = Ns - 1;
    constexpr int iters = (total + T - 1) / T;
#pragma unroll
    for (int it = 0; it < iters; ++it) {
        const int idx = it * T + threadIdx.x;
        if (total % T != 0 && idx >= total) break;
        const int lane = idx >> lgPer, j = idx & (per - 1);
        const int k = j & (Ns - 1);
        const int base_out = ((j - k) << lgR) + k;
        const float2 *xl = x + lane * PITCH;
        float2 *yl = y + lane * PITCH;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = xl[j + r * per];
            if (r > 0 && LGNS > 0) v[r] = cmulf(v[r], tw[tbase + (R - 1) * k + (r - 1)]);
        }
        if (R == 4) {
            bfly4(v[0], v[1], v[2], v[3]);
        } else {
            const float2 a0 = caddf(v[0], v[1]), a1 = csubf(v[0], v[1]);
            v[0] = a0; v[1] = a1;
        }
        if (LGNS == 0 && R == 4 && (PITCH % 2 == 0)) {        // first stage: 4 consecutive outputs, two 16-byte stores
            *reinterpret_cast<float4 *>(yl + base_out) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
            *reinterpret_cast<float4 *>(yl + base_out + 2) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) yl[base_out + (r << LGNS)] = v[r];
        }
    }
}

// barrier among the FFT threads: the whole CTA, or (NAMED) only the 256 consumer threads of the
// TMA-fed kernel, whose producer warp must not take part
template <bool NAMED>
__device__ __forceinline__ void fft_sync() {
    if (NAMED) asm volatile("bar.sync 1, 256;" ::: "memory");
    else __syncthreads();
}

template <int LGN, int LGNS, int LANES, int PITCH, int CUR, bool NAMED = false, int T = FftThreads<LGN>::value>
__device__ __forceinline__ int fft_ct(float2 *buf0, float2 *buf1, const float2 *tw) {
    fft_sync<NAMED>();
    if constexpr (LGNS + 2 <= LGN) {
        stage_ct<4, LGN, LGNS, LANES, PITCH, T>(CUR ? buf1 : buf0, CUR ? buf0 : buf1, tw);
        return fft_ct<LGN, LGNS + 2, LANES, PITCH, CUR ^ 1, NAMED, T>(buf0, buf1, tw);
    } else if constexpr (LGNS < LGN) {
        stage_ct<2, LGN, LGNS, LANES, PITCH, T>(CUR ? buf1 : buf0, CUR ? buf0 : buf1, tw);
        return fft_ct<LGN, LGNS + 1, LANES, PITCH, CUR ^ 1, NAMED, T>(buf0, buf1, tw);
    } else {
        return CUR;
    }
}

// ---- TMA-fed persistent row pass ------------------------------------------------------------------------
// The HBM-bound part of NF->FF as a producer/consumer pipeline: one producer warp streams the s1*s2 aliased
// row segments of each folded row into a ring of shared-memory slots with 1-D bulk copies (cp.async.bulk,
// mbarrier transaction counts); 256 consumer threads add the segments into registers as they land, then run
// the row FFT and store the row.  The producer keeps prefetching the next rows while the consumers are in
// their butterfly stages, so loads stay in flight all the time (the thread-issued loader loses ~20 % of the
// HBM bandwidth to that phase alternation).  CTAs are persistent and stride over the (field, row) work items
// (drawing them from a device counter was built in round 1 and measured slower in the pipelined step: removed);
// the producer hands the item number to the consumers through the slot of the item's first segment (-1 = no
// more work).
constexpr int TMA_CONSUMERS = 256;

template <int LGN, int RING_KB = 64>
struct RowsTmaCfg {
    static constexpr int N = 1 << LGN;
    static constexpr int SLOT_BYTES = N * 8;
    static constexpr int SLOTS = (RING_KB * 1024 / SLOT_BYTES) > 32 ? 32 : ((RING_KB * 1024 / SLOT_BYTES) < 4 ? 4 : (RING_KB * 1024 / SLOT_BYTES));
    static constexpr int EPT = N / TMA_CONSUMERS;                       // complex elements per consumer thread
    static constexpr size_t SMEM = (size_t)SLOTS * SLOT_BYTES + 3 * (size_t)N * 8 + 2 * SLOTS * sizeof(uint64_t) +
                                   SLOTS * sizeof(int) + 16;
};

template <int LGN, int RING_KB = 64>
__global__ void __launch_bounds__(TMA_CONSUMERS + 32, 1) fft_rows_tma_kernel(const FftArgs a, int batch,
                                                                              const __grid_constant__ RowScatter sc) {
    using Cfg = RowsTmaCfg<LGN, RING_KB>;
    constexpr int N = Cfg::N, S = Cfg::SLOTS, EPT = Cfg::EPT;
    extern __shared__ __align__(128) unsigned char tsm[];
    float2 *ring = reinterpret_cast<float2 *>(tsm);
    float2 *buf0 = ring + (size_t)S * N, *buf1 = buf0 + N, *stw = buf1 + N;
    uint64_t *full = reinterpret_cast<uint64_t *>(stw + N), *empty = full + S;
    int *wq = reinterpret_cast<int *>(empty + S);                           // work item of the segment in each slot
    const int tid = threadIdx.x;
    const int nseg = a.s1 * a.s2;
    const int work = a.other * batch;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], TMA_CONSUMERS / 32); }
        mbar_fence_init();
    }
    for (int t = tid; t < N; t += blockDim.x) stw[t] = a.tw[N + t];          // staged twiddle table
    __syncthreads();

    if (tid >= TMA_CONSUMERS) {
        // ------------------------------------------------ producer warp (one lane issues)
        if (tid == TMA_CONSUMERS) {
            long long g = 0;
            const uint64_t pol = l2_evict_first_policy();
            for (int seq = 0;; ++seq) {
                int w = blockIdx.x + seq * (int)gridDim.x;
                if (w >= work) w = -1;
                const int f = w < 0 ? 0 : w / a.other, r = w < 0 ? 0 : w - f * a.other;
                int rs = r - a.in_roll_r; if (rs < 0) rs += a.other;
                const float2 *src = pick4(a.in, f);
                for (int t = 0; t < (w < 0 ? 1 : nseg); ++t, ++g) {
                    const int slot = (int)(g % S);
                    const long long use = g / S;
                    if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));
                    wq[slot] = w;                                           // released by the arrive below
                    if (w < 0) { mbar_arrive(&full[slot]); break; }         // end marker: completes the phase, no data
                    const int t1 = t / a.s2, t2 = t - t1 * a.s2;
                    mbar_expect_tx(&full[slot], Cfg::SLOT_BYTES);
                    const float2 *seg = src + (size_t)(rs + t1 * a.other) * a.ld_in + ((size_t)t2 << LGN);
                    if (a.evict_first) bulk_g2s_hint(ring + (size_t)slot * N, seg, Cfg::SLOT_BYTES, &full[slot], pol);
                    else bulk_g2s(ring + (size_t)slot * N, seg, Cfg::SLOT_BYTES, &full[slot]);
                }
                if (w < 0) break;
            }
        }
        return;
    }
    // ---------------------------------------------------- consumers
    long long g = 0;
    for (;;) {
        mbar_wait(&full[(int)(g % S)], (uint32_t)((g / S) & 1));             // first segment of the next item (or the end marker)
        const int w = wq[(int)(g % S)];
        if (w < 0) break;
        const int f = w / a.other, r = w - f * a.other;
        float acc[2 * EPT];
#pragma unroll
        for (int v = 0; v < 2 * EPT; ++v) acc[v] = 0.f;
        for (int t = 0; t < nseg; ++t, ++g) {
            const int slot = (int)(g % S);
            if (t > 0) mbar_wait(&full[slot], (uint32_t)((g / S) & 1));
            const float2 *sl = ring + (size_t)slot * N;
            if (EPT >= 2) {
#pragma unroll
                for (int v = 0; v < EPT / 2; ++v) {
                    const float4 x = *reinterpret_cast<const float4 *>(sl + (v * TMA_CONSUMERS + tid) * 2);
                    acc[4 * v] += x.x; acc[4 * v + 1] += x.y; acc[4 * v + 2] += x.z; acc[4 * v + 3] += x.w;
                }
            } else {
                const float2 x = sl[tid];
                acc[0] += x.x; acc[1] += x.y;
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[slot]);               // this warp is done with the slot
        }
        // rotate by the input fftshift while handing the folded row to the FFT buffer
        if (EPT >= 2) {
#pragma unroll
            for (int v = 0; v < EPT / 2; ++v) {
                const int c = (v * TMA_CONSUMERS + tid) * 2;
                buf0[(c + a.in_roll_c) & (N - 1)] = make_float2(acc[4 * v], acc[4 * v + 1]);
                buf0[(c + 1 + a.in_roll_c) & (N - 1)] = make_float2(acc[4 * v + 2], acc[4 * v + 3]);
            }
        } else {
            buf0[(tid + a.in_roll_c) & (N - 1)] = make_float2(acc[0], acc[1]);
        }
        const int cur = fft_ct<LGN, 0, 1, N, 0, true>(buf0, buf1, stw);
        const float2 *res = cur ? buf1 : buf0;
        if (sc.world) {
            // fused all-to-all: the row's column slabs go straight to their owners over NVLink (8-byte stores, a warp
            // covers 256 contiguous bytes of one peer's row), overlapped with the TMA stream of the next rows
            int R = sc.out_row0 + (r >> sc.lg_block) * sc.row_stride + (r & ((1 << sc.lg_block) - 1));
            if (R >= sc.rows_total) R -= sc.rows_total;
            const size_t row_off = (size_t)R * a.ld_out;
            const int slab_mask = (1 << sc.lg_slab) - 1;
#pragma unroll
            for (int v = 0; v < EPT; ++v) {
                const int n = v * TMA_CONSUMERS + tid;
                float2 *dst = sc.out[n >> sc.lg_slab][f];
                dst[row_off + (n & slab_mask)] = res[(n - a.out_roll) & (N - 1)];
            }
        } else if (a.transpose_out) {
            // out[n][r]: the next pass transforms along r, so it finds ITS rows contiguous.  8-byte stores one
            // row pitch apart; the 33 MB intermediate lives in L2, which merges them into full sectors.
            float2 *dst = pick4(a.out, f) + r;
#pragma unroll
            for (int v = 0; v < EPT; ++v) {
                const int n = v * TMA_CONSUMERS + tid;
                dst[(size_t)n * a.ld_out] = res[(n - a.out_roll) & (N - 1)];
            }
        } else {
            float2 *dst = pick4(a.out, f) + (size_t)r * a.ld_out;
#pragma unroll
            for (int v = 0; v < EPT; ++v) {
                const int n = v * TMA_CONSUMERS + tid;                        // position in the OUTPUT row
                dst[n] = res[(n - a.out_roll) & (N - 1)];
            }
        }
        fft_sync<true>();                                                     // buffers free for the next row
    }
}

// rows: `lanes` consecutive rows per CTA, transform along the contiguous axis.  The loader sums the
// s1 x s2 aliased copies (aperture fold), applies the input fftshift and performs the first radix-4
// stage in registers, so its shared-memory writes are contiguous 64-byte runs.
template <int VEC, int LGN>
__global__ void __launch_bounds__(FftThreads<LGN>::value, (LGN >= 8 && LGN <= 11) ? 8 : 1) fft_rows_kernel(const FftArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int N = 1 << a.lgN, L = a.lanes;
    float2 *buf0 = fsm, *buf1 = fsm + (size_t)L * N, *stw = buf1 + (size_t)L * N;
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    const int row0 = blockIdx.x * L;
    for (int t = threadIdx.x; t < N; t += blockDim.x) stw[t] = a.tw[(LGN > 0 ? N : 0) + t];   // (staged) twiddle table -> smem
    int lgNs0;
    if (a.plain_loader) {
        // plain loader: consecutive samples per thread, every FFT stage in shared memory
        lgNs0 = 0;
        const int total = (L << a.lgN) / VEC;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int e = idx * VEC;
            const int lane = e >> a.lgN, n = e & (N - 1);
            const int r = row0 + lane;
            float acc[2 * VEC];
            if (r < a.other) {
                int rs = r - a.in_roll_r; if (rs < 0) rs += a.other;
                load_folded<VEC>(a, in, rs, n, N, acc);
            } else {
#pragma unroll
                for (int v = 0; v < 2 * VEC; ++v) acc[v] = 0.f;
            }
            if (VEC == 2) *reinterpret_cast<float4 *>(buf0 + e) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            else buf0[e] = make_float2(acc[0], acc[1]);
        }
    } else if (a.lgN >= 2) {
        lgNs0 = 2;
        const int lgPer = a.lgN - 2, per = 1 << lgPer;
        const int total = (L << lgPer) / VEC;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int e = idx * VEC;
            const int lane = e >> lgPer, j = e & (per - 1);
            const int r = row0 + lane;
            float acc[4][2 * VEC];
            if (r < a.other) {
                int rs = r - a.in_roll_r; if (rs < 0) rs += a.other;
#pragma unroll
                for (int q = 0; q < 4; ++q) load_folded<VEC>(a, in, rs, j + q * per, N, acc[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int v = 0; v < 2 * VEC; ++v) acc[q][v] = 0.f;
            }
            float2 *y = buf0 + lane * N + 4 * j;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float2 v0 = make_float2(acc[0][2 * v], acc[0][2 * v + 1]), v1 = make_float2(acc[1][2 * v], acc[1][2 * v + 1]);
                float2 v2 = make_float2(acc[2][2 * v], acc[2][2 * v + 1]), v3 = make_float2(acc[3][2 * v], acc[3][2 * v + 1]);
                bfly4(v0, v1, v2, v3);
                *reinterpret_cast<float4 *>(y + 4 * v) = make_float4(v0.x, v0.y, v1.x, v1.y);
                *reinterpret_cast<float4 *>(y + 4 * v + 2) = make_float4(v2.x, v2.y, v3.x, v3.y);
            }
        }
    } else {                                                   // N == 2: the whole transform is one butterfly
        lgNs0 = 1;
        for (int lane = threadIdx.x; lane < L; lane += blockDim.x) {
            const int r = row0 + lane;
            float x0[2] = {0.f, 0.f}, x1[2] = {0.f, 0.f};
            if (r < a.other) {
                int rs = r - a.in_roll_r; if (rs < 0) rs += a.other;
                load_folded<1>(a, in, rs, 0, N, x0);
                load_folded<1>(a, in, rs, 1, N, x1);
            }
            buf0[lane * N] = make_float2(x0[0] + x1[0], x0[1] + x1[1]);
            buf0[lane * N + 1] = make_float2(x0[0] - x1[0], x0[1] - x1[1]);
        }
    }
    int cur;
    if constexpr (LGN > 0) {
        constexpr int CL = (1024 >> LGN) > 0 ? (1024 >> LGN) : 1;          // rows per CTA (host uses the same rule)
        cur = (lgNs0 == 0) ? fft_ct<LGN, 0, CL, (1 << LGN), 0>(buf0, buf1, stw) : fft_ct<LGN, 2, CL, (1 << LGN), 0>(buf0, buf1, stw);
    } else {
        cur = fft_in_smem(buf0, buf1, a.lgN, lgNs0, L, N, stw);
    }
    const float2 *res = cur ? buf1 : buf0;
    const int tot = L << a.lgN;
    for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        const int lane = idx >> a.lgN, n = idx & (N - 1);       // n = position in the OUTPUT row
        const int r = row0 + lane;
        if (r < a.other) {
            const int q = (n - a.out_roll) & (N - 1);           // out[(q + roll) % N] = X[q]
            out[(size_t)r * a.ld_out + n] = res[lane * N + q];
        }
    }
}

// columns: `lanes` adjacent columns per CTA, transform along the strided axis.  Columns are
// transposed into [lane][n] shared-memory rows (pitch N+4) while the loader does the first stage.
template <int LGN, int CL>
__global__ void __launch_bounds__(FftThreads<LGN>::value) fft_cols_kernel(const FftArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int N = 1 << a.lgN, L = a.lanes, P = N + 4;
    float2 *buf0 = fsm, *buf1 = fsm + (size_t)L * P, *stw = buf1 + (size_t)L * P;
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    const int c0 = blockIdx.x * L;
    for (int t = threadIdx.x; t < N; t += blockDim.x) stw[t] = a.tw[(LGN > 0 ? N : 0) + t];
    int lgNs0 = 2;
    if constexpr (LGN > 0) {
        // compile-time sized: every load of the thread is issued before the first butterfly (16 loads in
        // flight per thread at 1024 x 4), lane/row splits are shifts
        constexpr int CN = 1 << LGN, per = CN >> 2, total = per * CL, T = FftThreads<LGN>::value;
        constexpr int iters = (total + T - 1) / T;
        constexpr int lgCL = (CL == 16) ? 4 : (CL == 8) ? 3 : (CL == 4) ? 2 : (CL == 2) ? 1 : 0;
        float2 v[iters][4];
        const size_t step = (size_t)per * a.ld_in;
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int idx = it * T + threadIdx.x;
            const int j = idx >> lgCL, lane = idx & (CL - 1);
            const int c = c0 + lane;
            if (idx < total && c < a.other) {
                const float2 *src = in + (size_t)j * a.ld_in + c;
                v[it][0] = src[0]; v[it][1] = src[step]; v[it][2] = src[2 * step]; v[it][3] = src[3 * step];
            } else {
                v[it][0] = v[it][1] = v[it][2] = v[it][3] = make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int it = 0; it < iters; ++it) {
            const int idx = it * T + threadIdx.x;
            if (idx < total) {
                const int j = idx >> lgCL, lane = idx & (CL - 1);
                bfly4(v[it][0], v[it][1], v[it][2], v[it][3]);
                float2 *y = buf0 + lane * P + 4 * j;
                *reinterpret_cast<float4 *>(y) = make_float4(v[it][0].x, v[it][0].y, v[it][1].x, v[it][1].y);
                *reinterpret_cast<float4 *>(y + 2) = make_float4(v[it][2].x, v[it][2].y, v[it][3].x, v[it][3].y);
            }
        }
    } else if (a.lgN >= 2) {
        const int per = N >> 2;
        const int total = per * L;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int j = idx / L, lane = idx - j * L;
            const int c = c0 + lane;
            float2 v0, v1, v2, v3;
            if (c < a.other) {
                const float2 *src = in + (size_t)j * a.ld_in + c;
                const size_t step = (size_t)per * a.ld_in;
                v0 = src[0]; v1 = src[step]; v2 = src[2 * step]; v3 = src[3 * step];
            } else {
                v0 = v1 = v2 = v3 = make_float2(0.f, 0.f);
            }
            bfly4(v0, v1, v2, v3);
            float2 *y = buf0 + lane * P + 4 * j;
            *reinterpret_cast<float4 *>(y) = make_float4(v0.x, v0.y, v1.x, v1.y);
            *reinterpret_cast<float4 *>(y + 2) = make_float4(v2.x, v2.y, v3.x, v3.y);
        }
    } else {
        lgNs0 = 1;
        for (int lane = threadIdx.x; lane < L; lane += blockDim.x) {
            const int c = c0 + lane;
            float2 x0 = make_float2(0.f, 0.f), x1 = x0;
            if (c < a.other) { x0 = in[c]; x1 = in[(size_t)a.ld_in + c]; }
            buf0[lane * P] = caddf(x0, x1);
            buf0[lane * P + 1] = csubf(x0, x1);
        }
    }
    int cur;
    if constexpr (LGN > 0) cur = fft_ct<LGN, 2, CL, (1 << LGN) + 4, 0>(buf0, buf1, stw);
    else cur = fft_in_smem(buf0, buf1, a.lgN, lgNs0, L, P, stw);
    const float2 *res = cur ? buf1 : buf0;
    if constexpr (LGN > 0) {
        constexpr int CN = 1 << LGN, total = CN * CL, T = FftThreads<LGN>::value, iters = total / T;
        constexpr int lgCL = (CL == 16) ? 4 : (CL == 8) ? 3 : (CL == 4) ? 2 : (CL == 2) ? 1 : 0;
#pragma unroll 8
        for (int it = 0; it < iters; ++it) {
            const int idx = it * T + threadIdx.x;
            const int n = idx >> lgCL, lane = idx & (CL - 1);       // n = OUTPUT row
            const int c = c0 + lane;
            if (c < a.other) out[(size_t)n * a.ld_out + c] = res[lane * P + ((n - a.out_roll) & (CN - 1))];
        }
    } else {
        const int total = L << a.lgN;
        for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int n = idx / L, lane = idx - n * L;              // n = OUTPUT row
            const int c = c0 + lane;
            if (c < a.other) {
                const int q = (n - a.out_roll) & (N - 1);
                out[(size_t)n * a.ld_out + c] = res[lane * P + q];
            }
        }
    }
}

// ---- fused column pass + radiated-power epilogue ---------------------------------------------------------
// The last two kernels of an FFT-path item in one: a CTA transforms CL adjacent columns of ALL FOUR fields
// (one field after the other through the same shared-memory ping-pong buffers) and keeps, per far-field point,
// only the two complex projections the radiated power needs,
//     t1 = L_phi + Z N_theta,   t2 = L_theta - Z N_phi          (nearfield_farfield.py:158-167, :184)
// which are real-coefficient combinations of the four aperture sums:
//     t1 = -px Fex - py Fey + Z cy Fhx - Z cx Fhy,   t2 = -cy Fex + cx Fey - Z px Fhx - Z py Fhy
// with (px,py) = (ux,uy)/(sin(theta)+1e-9), (cx,cy) = (px,py) uz; the DC bin (:161-169) is (px,py,cx,cy) =
// (1,0,1,0).  So the K x K x 4 aperture sums never go to memory (optional: fhat != NULL stores them), and
// P = k^2/(32 pi^2 Z) (|t1|^2+|t2|^2)/(uz+1e-5) * 2 (:184-189) is written straight from registers.
// Same arithmetic as ff_epilogue_kernel<true>: float64 evanescent mask / DC test, fp32 projections.
struct ColsPowerArgs {
    const float2 *in[4];
    float2 *fhat[4];
    const float2 *tw;
    const double *ux, *uy;
    float *P;
    double *block_sums;
    double scale;                  // pref * amp_scale^2 * 2
    float Z;
    int ld_in, ldf, ldp, n_cols, out_roll, accumulate, store_fhat;
};

template <int LGN, int CL, int T>
__global__ void __launch_bounds__(T, (T <= 512) ? 2 : 1) fft_cols_power_kernel(const ColsPowerArgs a) {
    constexpr int N = 1 << LGN, PITCH = N + 4, per = N >> 2, ltotal = per * CL, liters = ltotal / T;
    constexpr int total = N * CL, EPT = total / T;
    constexpr int lgCL = (CL == 16) ? 4 : (CL == 8) ? 3 : (CL == 4) ? 2 : (CL == 2) ? 1 : 0;
    static_assert(total % T == 0 && ltotal % T == 0 && T % CL == 0, "tile must divide evenly over the threads");
    extern __shared__ __align__(16) float2 fsm[];
    float2 *buf0 = fsm, *buf1 = fsm + CL * PITCH, *stw = buf1 + CL * PITCH;
    const int tid = threadIdx.x, c0 = blockIdx.x * CL;
    for (int t = tid; t < N; t += T) stw[t] = a.tw[N + t];
    // every element of this thread sits in the same column (T % CL == 0)
    const int lane = tid & (CL - 1), c = c0 + lane;
    const bool col_ok = c < a.n_cols;
    const double uy = col_ok ? a.uy[c] : 0.0;
    const double uy2 = __dmul_rn(uy, uy);
    // per point only 1/(sin(theta)+1e-9) and uz stay in registers; (float)ux comes from a shared-memory table
    float *uxf = reinterpret_cast<float *>(stw + N);
    for (int t = tid; t < N; t += T) uxf[t] = (float)a.ux[t];
    const float uyf = (float)uy;
    float inv[EPT], uzf[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int n = (e * T + tid) >> lgCL;                       // far-field row (ux index)
        const double ux = a.ux[n];
        // uz^2 exactly as numpy evaluates (1 - ux**2 - uy**2): bit-identical evanescent mask (:153-155)
        const double ux2 = __dmul_rn(ux, ux);
        const double uz2 = __dsub_rn(__dsub_rn(1.0, ux2), uy2);
        uzf[e] = (uz2 < 0.0) ? CUDART_NAN_F : sqrtf((float)uz2);
        inv[e] = 1.0f / ((float)sqrt(__dadd_rn(ux2, uy2)) + 1e-9f);
    }
    float2 t1[EPT], t2[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) t1[e] = t2[e] = make_float2(0.f, 0.f);
    const float Z = a.Z;
#pragma unroll 1
    for (int f = 0; f < 4; ++f) {
        const float2 *__restrict__ in = pick4(a.in, f);
        float2 *__restrict__ fh = pick4(a.fhat, f);
        float2 v[liters][4];
        const size_t step = (size_t)per * a.ld_in;
#pragma unroll
        for (int it = 0; it < liters; ++it) {
            const int idx = it * T + tid;
            const int j = idx >> lgCL;
            if (col_ok) {
                const float2 *src = in + (size_t)j * a.ld_in + c;
                v[it][0] = src[0]; v[it][1] = src[step]; v[it][2] = src[2 * step]; v[it][3] = src[3 * step];
            } else {
                v[it][0] = v[it][1] = v[it][2] = v[it][3] = make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int it = 0; it < liters; ++it) {
            const int j = (it * T + tid) >> lgCL;
            bfly4(v[it][0], v[it][1], v[it][2], v[it][3]);
            float2 *y = buf0 + lane * PITCH + 4 * j;
            *reinterpret_cast<float4 *>(y) = make_float4(v[it][0].x, v[it][0].y, v[it][1].x, v[it][1].y);
            *reinterpret_cast<float4 *>(y + 2) = make_float4(v[it][2].x, v[it][2].y, v[it][3].x, v[it][3].y);
        }
        const int cur = fft_ct<LGN, 2, CL, PITCH, 0, false, T>(buf0, buf1, stw);
        const float2 *res = (cur ? buf1 : buf0) + lane * PITCH;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int n = (e * T + tid) >> lgCL;
            const float2 w = res[(n - a.out_roll) & (N - 1)];
            if (a.store_fhat && col_ok) fh[(size_t)n * a.ldf + c] = w;
            const float uxn = uxf[n];
            const bool dc = (uxn == 0.f) && (uyf == 0.f);          // (float)u == 0 iff u == 0 on these grids
            const float pxe = dc ? 1.f : uxn * inv[e], pye = uyf * inv[e];
            const float cx = pxe * uzf[e], cy = pye * uzf[e];
            float k1, k2;
            if (f == 0) { k1 = -pxe; k2 = -cy; }                   // Ex:  L_y = -Fex
            else if (f == 1) { k1 = -pye; k2 = cx; }               // Ey:  L_x =  Fey
            else if (f == 2) { k1 = Z * cy; k2 = -(Z * pxe); }     // Hx:  N_y =  Fhx
            else { k1 = -(Z * cx); k2 = -(Z * pye); }              // Hy:  N_x = -Fhy
            t1[e].x = fmaf(k1, w.x, t1[e].x); t1[e].y = fmaf(k1, w.y, t1[e].y);
            t2[e].x = fmaf(k2, w.x, t2[e].x); t2[e].y = fmaf(k2, w.y, t2[e].y);
        }
        __syncthreads();                                           // buffers free for the next field
    }
    double sum = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
        const int n = (e * T + tid) >> lgCL;
        const float mag = t1[e].x * t1[e].x + t1[e].y * t1[e].y + t2[e].x * t2[e].x + t2[e].y * t2[e].y;
        const double p = a.scale * (double)mag / ((double)uzf[e] + 1e-5);
        if (col_ok) {
            float pf = (float)p;
            float *dst = a.P + (size_t)n * a.ldp + c;
            if (a.accumulate) pf += *dst;                          // incoherent sum over sources
            *dst = pf;
            if (isfinite(pf)) sum += p;
        }
    }
    if (a.block_sums) {                                            // :74 total_P over finite bins
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __shared__ double ws[T / 32];
        if ((tid & 31) == 0) ws[tid >> 5] = sum;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < T / 32; ++w) s += ws[w];
            a.block_sums[blockIdx.x] = s;
        }
    }
}

// ---- long column transforms (N >= 4096): four-step decomposition N = A * B ------------------------------
// A strided column of 4096/8192 points does not fit a shared-memory tile that is also wide enough for
// coalesced 128-byte row segments, so the column pass becomes two passes over 16-column tiles:
//   pass 1, for every n2 < B:  A-point DFT over rows n2 + n1*B, times W_N^(n2*k1), stored in place
//   pass 2, for every k1 < A:  B-point DFT over rows k1*B + n2, stored at row (k1 + A*k2 + roll) mod N
// (X[k1 + A k2] = sum_n2 W_B^(n2 k2) W_N^(n2 k1) sum_n1 x[B n1 + n2] W_A^(n1 k1)).  Both passes move
// 128-byte row segments; the sub-transforms reuse the compile-time Stockham stages.
struct FftSubArgs {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;          // plain table W_N^t of the FULL length N (first N entries of mlb_fft_twiddle)
    int ld_in, ld_out, lgNtot, n_cols;
    int in_gs, in_rs;          // input row  = g*in_gs  + n*in_rs
    int out_gs, out_rs, roll;  // output row = (g*out_gs + q*out_rs + roll) mod N
    int twiddle;               // multiply output q of sub-transform g by W_N^(g*q)
};

template <int LGA, int CL>
__global__ void __launch_bounds__(256) fft_cols_sub_kernel(const FftSubArgs a) {
    constexpr int A = 1 << LGA, P = A + 4, per = A >> 2, total = per * CL, iters = (total + 255) / 256;
    constexpr int lgCL = (CL == 16) ? 4 : (CL == 8) ? 3 : (CL == 4) ? 2 : (CL == 2) ? 1 : 0;
    __shared__ __align__(16) float2 buf0[CL * P];
    __shared__ __align__(16) float2 buf1[CL * P];
    __shared__ float2 stw[A];
    const float2 *__restrict__ in = pick4(a.in, blockIdx.z);
    float2 *__restrict__ out = pick4(a.out, blockIdx.z);
    const int c0 = blockIdx.x * CL, g = blockIdx.y;
    const int Nmask = (1 << a.lgNtot) - 1;
    // staged twiddle table of the A-point sub-transform, taken from the length-N table: W_A^t = W_N^(t N/A)
    for (int t = threadIdx.x; t < A; t += 256) {
        float2 w = make_float2(1.f, 0.f);
        int Ns = 1, lg = 0;
        bool done = false;
        while (lg + 2 <= LGA && !done) {
            if (t >= Ns - 1 && t < 4 * Ns - 1) {
                const int e = t - (Ns - 1), k = e / 3, r = e - 3 * k + 1;
                w = __ldg(a.tw + (((r * k) << (a.lgNtot - lg - 2)) & Nmask));
                done = true;
            }
            Ns *= 4; lg += 2;
        }
        if (!done && lg < LGA && t >= Ns - 1 && t < 2 * Ns - 1) w = __ldg(a.tw + (((t - (Ns - 1)) << (a.lgNtot - lg - 1)) & Nmask));
        stw[t] = w;
    }
    float2 v[iters][4];
#pragma unroll
    for (int it = 0; it < iters; ++it) {
        const int idx = it * 256 + threadIdx.x;
        const int j = idx >> lgCL, lane = idx & (CL - 1);
        const int c = c0 + lane;
        if (idx < total && c < a.n_cols) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                v[it][r] = in[(size_t)(g * a.in_gs + (j + r * per) * a.in_rs) * a.ld_in + c];
        } else {
            v[it][0] = v[it][1] = v[it][2] = v[it][3] = make_float2(0.f, 0.f);
        }
    }
#pragma unroll
    for (int it = 0; it < iters; ++it) {
        const int idx = it * 256 + threadIdx.x;
        if (idx < total) {
            const int j = idx >> lgCL, lane = idx & (CL - 1);
            bfly4(v[it][0], v[it][1], v[it][2], v[it][3]);
            float2 *y = buf0 + lane * P + 4 * j;
            *reinterpret_cast<float4 *>(y) = make_float4(v[it][0].x, v[it][0].y, v[it][1].x, v[it][1].y);
            *reinterpret_cast<float4 *>(y + 2) = make_float4(v[it][2].x, v[it][2].y, v[it][3].x, v[it][3].y);
        }
    }
    const int cur = fft_ct<LGA, 2, CL, P, 0>(buf0, buf1, stw);
    const float2 *res = cur ? buf1 : buf0;
    constexpr int otot = A * CL, oit = (otot + 255) / 256;
#pragma unroll 4
    for (int it = 0; it < oit; ++it) {
        const int idx = it * 256 + threadIdx.x;
        const int q = idx >> lgCL, lane = idx & (CL - 1);
        const int c = c0 + lane;
        if (idx < otot && c < a.n_cols) {
            float2 w = res[lane * P + q];
            if (a.twiddle) w = cmulf(w, __ldg(a.tw + ((g * q) & Nmask)));
            const int orow = (g * a.out_gs + q * a.out_rs + a.roll) & Nmask;
            out[(size_t)orow * a.ld_out + c] = w;
        }
    }
}

// ---- small in-register DFTs shared with the big-radix mixed engine (fftmix.cuh) ----------------------------------
template <int R>
__device__ __forceinline__ void dft_small(float2 (&v)[R]) {
    if (R == 2) {
        const float2 a = caddf(v[0], v[1]), b = csubf(v[0], v[1]);
        v[0] = a; v[1] = b;
    } else if (R == 4) {
        bfly4(v[0], v[1], v[2], v[3]);
    } else if (R == 3) {
        const float s3 = 0.86602540378443864676f;                      // sin(2 pi / 3)
        const float2 t1 = caddf(v[1], v[2]), t2 = csubf(v[1], v[2]);
        const float2 m = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
        const float2 r = make_float2(s3 * t2.y, -s3 * t2.x);          // -i * s3 * t2
        v[0] = caddf(v[0], t1);
        v[1] = caddf(m, r);
        v[2] = csubf(m, r);
    } else {                                                           // R == 5
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;   // cos(2pi/5), cos(4pi/5)
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;    // sin(2pi/5), sin(4pi/5)
        const float2 a1 = caddf(v[1], v[4]), b1 = csubf(v[1], v[4]);
        const float2 a2 = caddf(v[2], v[3]), b2 = csubf(v[2], v[3]);
        const float2 m1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
        const float2 m2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
        // -i * (s1 b1 + s2 b2)  and  -i * (s2 b1 - s1 b2)
        const float2 n1 = make_float2(s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x));
        const float2 n2 = make_float2(s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x));
        v[0] = caddf(v[0], caddf(a1, a2));
        v[1] = caddf(m1, n1);
        v[4] = csubf(m1, n1);
        v[2] = caddf(m2, n2);
        v[3] = csubf(m2, n2);
    }
}

}  // namespace mlb
#include "fftmix.cuh"
namespace mlb {

__global__ void fft_twiddle_kernel(int N, float2 *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    double s, c;
    sincospi(-2.0 * (double)t / (double)N, &s, &c);
    out[t] = make_float2((float)c, (float)s);                 // plain table W_N^t
    // staged table entry t (see stage_ct): find the stage whose block contains t
    float2 w = make_float2(1.f, 0.f);
    int Ns = 1, lg = 0, lgN = 0;
    while ((1 << lgN) < N) ++lgN;
    bool done = false;
    while (lg + 2 <= lgN && !done) {                          // radix-4 stages: block [Ns-1, 4Ns-1)
        if (t >= Ns - 1 && t < 4 * Ns - 1) {
            const int e = t - (Ns - 1), k = e / 3, r = e - 3 * k + 1;
            sincospi(-2.0 * (double)(r * k) / (double)(4 * Ns), &s, &c);
            w = make_float2((float)c, (float)s);
            done = true;
        }
        Ns *= 4; lg += 2;
    }
    if (!done && lg < lgN && t >= Ns - 1 && t < 2 * Ns - 1) { // final radix-2 stage: block [Ns-1, 2Ns-1)
        const int k = t - (Ns - 1);
        sincospi(-2.0 * (double)k / (double)(2 * Ns), &s, &c);
        w = make_float2((float)c, (float)s);
    }
    out[N + t] = w;
}

// tuning knobs (mlb_fft_tune): rows-pass loader variant, lanes and threads; defaults chosen on B200
static int g_rows_per_sm = 0;          // TMA row pass: resident CTAs per SM (0 = as many as fit, at most 3)
static int g_rows_engine = 2;          // row pass: 0 = radix-4 shared-memory kernels (TMA-fed where possible), 1 = radix-16
                                       // register kernels (256..8192 points), 2 = radix-16 only without a fold (default:
                                       // the TMA-fed fold+FFT kernel stays the choice when the aperture is folded)
static int g_r16_occ = 0;              // radix-16 kernels: resident CTAs per SM they are compiled for (0 = default, 2..4)
static int g_cols_engine = 1;          // column pass: 0 = radix-4 shared-memory kernels, 1 = radix-16 register kernels (256..8192; default)
static int g_rows_ring_kb = 64;        // TMA row pass: bytes of shared memory in the slot ring per CTA (64 or 128 KB)
static int g_rows_evict_first = 1;     // TMA row pass: stream the aperture through L2 with an evict-first policy
static int g_cols_strip_mb = 0;        // two-pass column transforms (>= 4096 points): run them strip by strip, strips of this many MB (so the
                                       // intermediate stays in L2); 0 = one strip (default: measured faster on B200, the per-strip launches cost
                                       // more than the saved HBM round trip)
static int g_r16_min_lg = 10;          // radix-16 kernels from 2^this points up (8..13): below 1024 points the radix-4 kernels launch more
                                       // CTAs and measure faster on B200 (cfg2 / the 256-point sweep of bench.py)
static int g_mixed_occ = 0;            // register kernels of the big-radix engine: resident CTAs per SM they are compiled for
                                       // (2..4; 0 = default: rows 2, columns 4 -- measured on B200, scripts/allbins_kernels.py)
static int g_mixed_reg = 1;            // big-radix engine: use the one-butterfly-per-thread register kernels where they apply
static int g_mixed_ct = 1;             // register kernels of the big-radix engine: 1 = the instantiation compiled for the plan (default), 0 = generic
static int g_cols_power_wide = -1;     // fused column+power pass: 1 = 4096-point tiles / 1024 threads, 0 = 2048 / 512,
                                       // -1 = by length (wide from 1024 points up: measured faster on B200)
static int g_rows_plain = 1, g_rows_points = 1024, g_rows_threads = 256, g_rows_vec = 2, g_rows_tma = 1, g_cols_half = 0;

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
// radix sequence (4s first, then 2, 3s, 5s) of a 5-smooth length; returns the stage count or 0
static int factor_235(int n, int *radix) {
    int ns = 0;
    while (n % 4 == 0) { radix[ns++] = 4; n /= 4; }
    while (n % 2 == 0) { radix[ns++] = 2; n /= 2; }
    while (n % 3 == 0) { radix[ns++] = 3; n /= 3; }
    while (n % 5 == 0) { radix[ns++] = 5; n /= 5; }
    return (n == 1 && ns <= 16) ? ns : 0;
}
static int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }
constexpr int FFT_MAX_N = 8192;                     // 2 x 8192 x 8 B = 128 KB of shared memory

static int fill_args(FftArgs &a, const mlb_c64 *const *h_in, mlb_c64 *const *h_out, int batch, const char *who) {
    MLB_REQUIRE(h_in && h_out && batch >= 1 && batch <= 4, "%s: bad batch %d", who, batch);
    for (int b = 0; b < 4; ++b) {
        const int s = b < batch ? b : 0;
        MLB_REQUIRE(h_in[s] && h_out[s], "%s: NULL operand %d", who, s);
        a.in[b] = reinterpret_cast<const float2 *>(h_in[s]);
        a.out[b] = reinterpret_cast<float2 *>(h_out[s]);
    }
    return MLB_OK;
}

}  // namespace mlb

extern "C" int mlb_fft_twiddle(int N, mlb_c64 *out, void *stream) {
    MLB_REQUIRE(N >= 1 && out, "mlb_fft_twiddle: bad arguments");
    mlb::fft_twiddle_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, reinterpret_cast<float2 *>(out));
    return mlb::check_launch("mlb_fft_twiddle");
}

extern "C" int mlb_fft_tune(int rows_plain_loader, int rows_points_per_cta, int rows_threads, int rows_vec) {
    // rows_plain_loader: 2 = TMA-fed persistent kernel (default), 1 = thread-issued loads, 0 = first stage in loader
    mlb::g_rows_tma = (rows_plain_loader == 2 || rows_plain_loader == 3);
    mlb::g_cols_half = (rows_plain_loader == 3);      // 3 = TMA rows + half-width column tiles
    if (rows_plain_loader >= 2) rows_plain_loader = 1;
    MLB_REQUIRE((rows_threads == 64 || rows_threads == 128 || rows_threads == 256) && rows_points_per_cta >= 1 &&
                    (rows_vec == 1 || rows_vec == 2),
                "mlb_fft_tune: bad arguments");
    mlb::g_rows_plain = rows_plain_loader ? 1 : 0;
    mlb::g_rows_points = rows_points_per_cta;
    mlb::g_rows_threads = rows_threads;
    mlb::g_rows_vec = rows_vec;
    return MLB_OK;
}

extern "C" int mlb_set_option(const char *name, int value) {
    MLB_REQUIRE(name != nullptr, "mlb_set_option: NULL name");
    const std::string n(name);
    if (n == "rows_ctas_per_sm") { MLB_REQUIRE(value >= 0 && value <= 3, "rows_ctas_per_sm: 0..3"); mlb::g_rows_per_sm = value; }
    else if (n == "rows_l2_evict_first") mlb::g_rows_evict_first = value ? 1 : 0;
    else if (n == "rows_engine") { MLB_REQUIRE(value >= 0 && value <= 2, "rows_engine: 0..2"); mlb::g_rows_engine = value; }
    else if (n == "r16_occupancy") { MLB_REQUIRE(value == 0 || (value >= 2 && value <= 4), "r16_occupancy: 0, 2..4"); mlb::g_r16_occ = value; }
    else if (n == "r16_min_lg") { MLB_REQUIRE(value >= 8 && value <= 13, "r16_min_lg: 8..13"); mlb::g_r16_min_lg = value; }
    else if (n == "cols_engine") { MLB_REQUIRE(value >= 0 && value <= 1, "cols_engine: 0..1"); mlb::g_cols_engine = value; }
    else if (n == "rows_ring_kb") { MLB_REQUIRE(value == 64 || value == 128, "rows_ring_kb: 64 or 128"); mlb::g_rows_ring_kb = value; }
    else if (n == "cols_power_wide") mlb::g_cols_power_wide = value < 0 ? -1 : (value ? 1 : 0);
    else if (n == "mixed_registers") { MLB_REQUIRE(value >= 0 && value <= 2, "mixed_registers: 0..2"); mlb::g_mixed_reg = value; }
    else if (n == "mixed_compiled") mlb::g_mixed_ct = value ? 1 : 0;
    else if (n == "mixed_occupancy") { MLB_REQUIRE(value == 0 || (value >= 2 && value <= 4), "mixed_occupancy: 0, 2..4"); mlb::g_mixed_occ = value; }
    else if (n == "cols_strip_mb") { MLB_REQUIRE(value >= 0 && value <= 4096, "cols_strip_mb: 0..4096"); mlb::g_cols_strip_mb = value; }
    else MLB_REQUIRE(false, "mlb_set_option: unknown option '%s'", name);
    return MLB_OK;
}

extern "C" int mlb_get_option(const char *name) {
    if (!name) return -1;
    const std::string n(name);
    if (n == "rows_ctas_per_sm") return mlb::g_rows_per_sm;
    if (n == "rows_l2_evict_first") return mlb::g_rows_evict_first;
    if (n == "rows_ring_kb") return mlb::g_rows_ring_kb;
    if (n == "rows_engine") return mlb::g_rows_engine;
    if (n == "cols_engine") return mlb::g_cols_engine;
    if (n == "r16_min_lg") return mlb::g_r16_min_lg;
    if (n == "r16_occupancy") return mlb::g_r16_occ;
    if (n == "cols_power_wide") return mlb::g_cols_power_wide;
    if (n == "cols_strip_mb") return mlb::g_cols_strip_mb;
    if (n == "mixed_registers") return mlb::g_mixed_reg;
    if (n == "mixed_occupancy") return mlb::g_mixed_occ;
    if (n == "mixed_compiled") return mlb::g_mixed_ct;
    return -1;
}

namespace mlb {
// ---- host side of the big-radix mixed engine (fftmix.cuh)
static unsigned mix2_magic(int d) { return d <= 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned long long)d - 1) / (unsigned long long)d); }

// radix sequence of a 5-smooth length: odd radices first (their stride-R first-stage writes are conflict-free),
// then the even ones; fills radix / magic tables; returns the stage count or 0
static int mix2_plan(Mix2Args &m, int N) {
    int a = 0, b = 0, c = 0, n = N;
    while (n % 2 == 0) { ++a; n /= 2; }
    while (n % 3 == 0) { ++b; n /= 3; }
    while (n % 5 == 0) { ++c; n /= 5; }
    if (n != 1 || N < 2) return 0;
    int r[32], ns = 0;
    while (b >= 1 && c >= 1) { r[ns++] = 15; --b; --c; }
    while (b >= 2) { r[ns++] = 9; b -= 2; }
    int five_even = 0, three_even = 0;                    // odd primes to be paired with factors of two
    while (c >= 1) { if (a >= 1 + 0 && five_even < a) { ++five_even; } else { r[ns++] = 5; } --c; }
    // pair each deferred 5 with one factor 2 (radix 10)
    int tens = five_even; a -= tens;
    if (b == 1) { if (a >= 2) { three_even = 12; a -= 2; } else if (a >= 1) { three_even = 6; a -= 1; } else { r[ns++] = 3; } b = 0; }
    if (three_even) r[ns++] = three_even;
    for (int t = 0; t < tens; ++t) r[ns++] = 10;
    while (a >= 4) { r[ns++] = 16; a -= 4; }
    if (a == 3) r[ns++] = 8; else if (a == 2) r[ns++] = 4; else if (a == 1) r[ns++] = 2;
    if (ns > MIX2_MAX_STAGES) return 0;
    int Ns = 1;
    for (int s = 0; s < ns; ++s) {
        m.radix[s] = r[s];
        m.magic_ns[s] = mix2_magic(Ns);
        m.magic_per[s] = mix2_magic(N / r[s]);
        Ns *= r[s];
    }
    m.nstage = ns; m.N = N;
    m.pad_sh = (r[0] & 1) ? 30 : 4;
    return ns;
}
static size_t mix2_smem(int N, int lanes) { return 2 * (size_t)lanes * (mixpad(N, 4) + 1) * sizeof(float2); }
constexpr size_t MIX2_SMEM_MAX = 2 * (size_t)(FFT_MAX_N + (FFT_MAX_N >> 4) + 1) * sizeof(float2);   // one 8192-point transform
static int mix2_set_smem() {
    static unsigned long long set_ = 0; const unsigned long long devbit_ = mlb::device_bit();
    if (!(set_ & devbit_)) {
        MLB_CUDA(cudaFuncSetAttribute(mix2_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX));
        MLB_CUDA(cudaFuncSetAttribute(mix2_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX));
        MLB_CUDA(cudaFuncSetAttribute(mix2_reg_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX / 2));
        MLB_CUDA(cudaFuncSetAttribute(mix2_reg_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX / 2));
        MLB_CUDA(cudaFuncSetAttribute(mix2_reg_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX / 2));
        MLB_CUDA(cudaFuncSetAttribute(mix2_reg_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX / 2));
        MLB_CUDA(cudaFuncSetAttribute(mix2_reg_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX / 2));
        MLB_CUDA(cudaFuncSetAttribute(mix2_reg_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MIX2_SMEM_MAX / 2));
        set_ |= devbit_;
    }
    return MLB_OK;
}
// power-of-two column lanes per CTA for sub-length n: as many as fit in ~72 KB, at most 32
static int mix2_col_lanes(int n, int n_cols) {
    int lanes = 1;
    while (lanes < 32 && mix2_smem(n, lanes * 2) <= 72 * 1024 && lanes < n_cols) lanes *= 2;
    return lanes;
}
// butterflies of the busiest stage of one transform (the register variant needs lanes * this <= 256)
static int mix2_max_per(const Mix2Args &m) {
    int mx = 0;
    for (int s = 0; s < m.nstage; ++s) mx = (m.N / m.radix[s] > mx) ? m.N / m.radix[s] : mx;
    return mx;
}
// compile-time plans of the register kernels (fftmix.cuh): every plan the dispatch below sends there
typedef void (*Mix2Kernel)(const Mix2Args);
struct Mix2CtEntry { int r0, r1, r2; Mix2Kernel fn[2]; };
#define MIX2_CT_ROW(r0, r1, r2) {r0, r1, r2, {mix2_ct_kernel<false, 3, r0, r1, r2>, mix2_ct_kernel<false, 4, r0, r1, r2>}}
#define MIX2_CT_COL(r0, r1) {r0, r1, 1, {mix2_ct_kernel<true, 3, r0, r1, 1>, mix2_ct_kernel<true, 4, r0, r1, 1>}}
static const Mix2CtEntry mix2_ct_rows[] = {         // N = 2160, 2250, 2304, 2400, 2560, 2700, 2880, 3072, 3375, 3600, 3840
    MIX2_CT_ROW(15, 9, 16), MIX2_CT_ROW(15, 15, 10), MIX2_CT_ROW(9, 16, 16), MIX2_CT_ROW(15, 10, 16), MIX2_CT_ROW(10, 16, 16),
    MIX2_CT_ROW(15, 15, 12), MIX2_CT_ROW(15, 12, 16), MIX2_CT_ROW(12, 16, 16), MIX2_CT_ROW(15, 15, 15), MIX2_CT_ROW(15, 15, 16),
    MIX2_CT_ROW(15, 16, 16)};
static const Mix2CtEntry mix2_ct_cols[] = {         // B = 30 .. 240: the second pass of N = A B columns (and short direct columns)
    MIX2_CT_COL(15, 2), MIX2_CT_COL(15, 3), MIX2_CT_COL(15, 4), MIX2_CT_COL(15, 5), MIX2_CT_COL(15, 6), MIX2_CT_COL(15, 8),
    MIX2_CT_COL(15, 9), MIX2_CT_COL(9, 16), MIX2_CT_COL(15, 10), MIX2_CT_COL(10, 16), MIX2_CT_COL(15, 12), MIX2_CT_COL(12, 16),
    MIX2_CT_COL(15, 15), MIX2_CT_COL(15, 16)};
#undef MIX2_CT_ROW
#undef MIX2_CT_COL
// the instantiation for plan m (lanes, padding and radices must all match), or NULL; occ = resident CTAs it is compiled for
template <size_t NE>
static Mix2Kernel mix2_ct_find(const Mix2CtEntry (&tab)[NE], const Mix2Args &m, int lanes, int occ) {
    if (!g_mixed_ct || m.lanes != lanes || (m.nstage != 2 && m.nstage != 3)) return nullptr;
    const int r2 = (m.nstage == 3) ? m.radix[2] : 1;
    for (const Mix2CtEntry &e : tab)
        if (e.r0 == m.radix[0] && e.r1 == m.radix[1] && e.r2 == r2 && m.N == e.r0 * e.r1 * e.r2 &&
            m.pad_sh == ((e.r0 & 1) ? 30 : 4))
            return e.fn[occ == 3 ? 0 : 1];
    return nullptr;
}

template <int A>
static void mix2_launch_first(const Mix2FirstArgs &f, dim3 grid, cudaStream_t st) { mix2_cols_first_kernel<A><<<grid, 256, 0, st>>>(f); }
static int mix2_first(int A, const Mix2FirstArgs &f, int batch, cudaStream_t st) {
    dim3 grid((f.n_cols + 31) / 32, (f.B + 7) / 8, batch);
    switch (A) {
        case 16: mix2_launch_first<16>(f, grid, st); break;
        case 15: mix2_launch_first<15>(f, grid, st); break;
        case 12: mix2_launch_first<12>(f, grid, st); break;
        case 10: mix2_launch_first<10>(f, grid, st); break;
        case 9: mix2_launch_first<9>(f, grid, st); break;
        case 8: mix2_launch_first<8>(f, grid, st); break;
        case 6: mix2_launch_first<6>(f, grid, st); break;
        case 5: mix2_launch_first<5>(f, grid, st); break;
        case 4: mix2_launch_first<4>(f, grid, st); break;
        case 3: mix2_launch_first<3>(f, grid, st); break;
        case 2: mix2_launch_first<2>(f, grid, st); break;
        default: set_error("mixed column pass: no first-pass radix %d", A); return MLB_ERR_ARG;
    }
    return check_launch("mlb_fft_cols(mixed radix, first pass)");
}
}  // namespace mlb

/* host only: the radix plan of the big-radix mixed engine for length N (no GPU needed) */
extern "C" int mlb_fft_mixed_plan(int N, int *radix8, int *pad_shift) {
    MLB_REQUIRE(radix8 != nullptr, "mlb_fft_mixed_plan: NULL output");
    mlb::Mix2Args m;
    const int ns = mlb::mix2_plan(m, N);
    MLB_REQUIRE(ns > 0, "mlb_fft_mixed_plan: %d is not of the form 2^a 3^b 5^c (or needs more than %d stages)", N,
                mlb::MIX2_MAX_STAGES);
    for (int s = 0; s < mlb::MIX2_MAX_STAGES; ++s) radix8[s] = s < ns ? m.radix[s] : 0;
    if (pad_shift) *pad_shift = m.pad_sh;
    return ns;
}

/* host only: 1 if the register kernels have an instantiation compiled for the plan of length N (cols = 0: a row of N
 * points, one row per CTA; cols = 1: a column (sub-)transform of N points, 16 columns per CTA), else 0 */
extern "C" int mlb_fft_mixed_compiled(int N, int cols) {
    mlb::Mix2Args m;
    if (N < 2 || N > mlb::FFT_MAX_N || mlb::mix2_plan(m, N) <= 0) return 0;
    m.lanes = cols ? mlb::MIX2_CT_COL_LANES : 1;
    const mlb::Mix2Kernel k = cols ? mlb::mix2_ct_find(mlb::mix2_ct_cols, m, mlb::MIX2_CT_COL_LANES, 0)
                                   : mlb::mix2_ct_find(mlb::mix2_ct_rows, m, 1, 0);
    return k != nullptr ? 1 : 0;
}

extern "C" int mlb_fft_max_length(void) { return mlb::FFT_MAX_N; }

extern "C" int mlb_fft_rows_can_transpose(int N) {
    return (mlb::g_rows_tma && mlb::is_pow2(N) && N >= 256 && N <= 2048) ? 1 : 0;
}

static int fft_rows_impl(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int n_rows,
                         int N, int s1, int s2, const mlb_c64 *tw, int in_roll_r, int in_roll_c, int out_roll,
                         int transpose_out, int batch, void *stream, const mlb::RowScatter *scatter = nullptr) {
    mlb::FftArgs a;
    static const mlb::RowScatter no_scatter = {};
    const mlb::RowScatter &sc = scatter ? *scatter : no_scatter;
    if (int rc = mlb::fill_args(a, h_in, h_out, batch, "mlb_fft_rows")) return rc;
    int radix_probe[16];
    MLB_REQUIRE(N >= 2 && N <= mlb::FFT_MAX_N && mlb::factor_235(N, radix_probe) > 0,
                "mlb_fft_rows: length %d must be of the form 2^a 3^b 5^c and <= %d", N, mlb::FFT_MAX_N);
    MLB_REQUIRE(s1 >= 1 && s2 >= 1, "mlb_fft_rows: fold factors must be >= 1");
    MLB_REQUIRE(tw && n_rows > 0 && ld_in >= N * s2 && ld_out >= (transpose_out ? n_rows : (scatter ? N / sc.world : N)),
                "mlb_fft_rows: bad sizes");
    a.transpose_out = transpose_out ? 1 : 0;
    MLB_REQUIRE(in_roll_r >= 0 && in_roll_r < n_rows && in_roll_c >= 0 && in_roll_c < N && out_roll >= 0 && out_roll < N,
                "mlb_fft_rows: rolls out of range");
    for (int b = 0; b < batch; ++b)
        MLB_REQUIRE((in_roll_r == 0 && s1 == 1 && s2 == 1) || a.in[b] != a.out[b],
                    "mlb_fft_rows: in-place needs in_roll_r == 0 and no fold");
    a.tw = reinterpret_cast<const float2 *>(tw);
    if (scatter && !(mlb::is_pow2(N) && N >= 256 && N <= 2048 && mlb::g_rows_tma && ld_in % 2 == 0)) {
        mlb::set_error("mlb_fft_rows_scatter: needs the TMA-fed row kernel (power-of-two N in 256..2048, even pitch); N = %d", N);
        return MLB_ERR_UNSUPPORTED;
    }
    if (!mlb::is_pow2(N)) {                        // mixed radix (good_fft_number sizes)
        MLB_REQUIRE(!transpose_out, "mlb_fft_rows: transposed output needs a power-of-two length");
        mlb::Mix2Args m2;
        MLB_REQUIRE(mlb::mix2_plan(m2, N) > 0, "mlb_fft_rows: cannot factor %d", N);
        for (int b = 0; b < 4; ++b) { m2.in[b] = a.in[b]; m2.out[b] = a.out[b]; }
        m2.tw = a.tw; m2.ld_in = ld_in; m2.ld_out = ld_out; m2.other = n_rows; m2.tw_mul = 1; m2.Ntot = N;
        m2.in_roll_r = in_roll_r; m2.in_roll_c = in_roll_c; m2.out_roll = out_roll; m2.s1 = s1; m2.s2 = s2;
        m2.in_gs = m2.in_rs = m2.out_gs = m2.out_rs = m2.roll = 0;
        int lanes = 1;                           // rows per CTA: ~one butterfly per thread and stage, <= 64 KB
        while (lanes < 16 && mlb::mix2_smem(N, lanes * 2) <= 64 * 1024 && lanes * 2 <= n_rows && lanes * N < 4096) lanes *= 2;
        if (int rc = mlb::mix2_set_smem()) return rc;
        const int mp = mlb::mix2_max_per(m2);
        int rl = mp <= 256 ? 256 / mp : 0;
        if (rl > 16) rl = 16;
        if (rl > n_rows) rl = n_rows;
        // one butterfly per thread and stage (register variant): long rows that keep >= 80 % of the threads busy
        if (mlb::g_mixed_reg && rl >= 1 && (mlb::g_mixed_reg == 2 || (N >= 2048 && rl * mp >= 205))) {
            m2.lanes = rl; m2.lg_lanes = 0;
            dim3 grid((n_rows + rl - 1) / rl, batch);
            const size_t sm2 = mlb::mix2_smem(N, rl) / 2;
            if (mlb::Mix2Kernel ct = (sm2 <= 48 * 1024) ? mlb::mix2_ct_find(mlb::mix2_ct_rows, m2, 1, mlb::g_mixed_occ) : nullptr) {
                ct<<<grid, 256, sm2, (cudaStream_t)stream>>>(m2);
                return mlb::check_launch("mlb_fft_rows(mixed radix, compiled plan)");
            }
            if (mlb::g_mixed_occ == 4) mlb::mix2_reg_kernel<false, 4><<<grid, 256, sm2, (cudaStream_t)stream>>>(m2);
            else if (mlb::g_mixed_occ == 3) mlb::mix2_reg_kernel<false, 3><<<grid, 256, sm2, (cudaStream_t)stream>>>(m2);
            else mlb::mix2_reg_kernel<false, 2><<<grid, 256, sm2, (cudaStream_t)stream>>>(m2);
            return mlb::check_launch("mlb_fft_rows(mixed radix, registers)");
        }
        m2.lanes = lanes; m2.lg_lanes = 0;
        dim3 grid((n_rows + lanes - 1) / lanes, batch);
        mlb::mix2_rows_kernel<<<grid, 256, mlb::mix2_smem(N, lanes), (cudaStream_t)stream>>>(m2);
        return mlb::check_launch("mlb_fft_rows(mixed radix)");
    }
    a.ld_in = ld_in; a.ld_out = ld_out; a.lgN = mlb::ilog2(N); a.other = n_rows;
    a.in_roll_r = in_roll_r; a.in_roll_c = in_roll_c; a.out_roll = out_roll; a.s1 = s1; a.s2 = s2;
    a.evict_first = mlb::g_rows_evict_first;
    int lanes = 1;
    while (lanes * 2 * N <= mlb::g_rows_points && lanes * 2 <= n_rows) lanes *= 2;   // small transforms: several rows per CTA
    a.lanes = lanes;
    a.plain_loader = mlb::g_rows_plain;
    const size_t smem = (2 * (size_t)lanes + 1) * N * sizeof(float2);     // ping-pong buffers + twiddle table
    bool vec = (mlb::g_rows_vec == 2) && (N >= 8) && (in_roll_c % 2 == 0) && (ld_in % 2 == 0);
    for (int b = 0; b < batch; ++b) vec = vec && mlb::aligned16(a.in[b]);
    dim3 grid((n_rows + lanes - 1) / lanes, batch);
    // radix-16 register kernels (256..8192 points)
    if (!scatter && !transpose_out && a.lgN >= mlb::g_r16_min_lg && a.lgN <= 13 &&
        (mlb::g_rows_engine == 1 || (mlb::g_rows_engine == 2 && s1 == 1 && s2 == 1))) {
        bool inplace = false;
        for (int b = 0; b < batch; ++b) inplace = inplace || (a.in[b] == a.out[b]);
        mlb::R16Args r;
        for (int b = 0; b < 4; ++b) { r.in[b] = a.in[b]; r.out[b] = a.out[b]; }
        r.tw = a.tw; r.ld_in = ld_in; r.ld_out = ld_out; r.n_rows = n_rows; r.in_roll_r = in_roll_r;
        r.in_roll_c = in_roll_c; r.out_roll = out_roll; r.s1 = s1; r.s2 = s2;
        cudaStream_t st = (cudaStream_t)stream;
        (void)inplace;      // in-place is safe: a CTA reads all of its rows before it writes them (in_roll_r == 0 checked above)
#define MLB_R16_ROWS(LG)                                                                                              \
    case LG: {                                                                                                       \
        constexpr int T = mlb::R16Threads<LG>::value, TR = (1 << LG) >> 4, L = T / TR;                               \
        constexpr int PITCH = (1 << LG) + ((1 << LG) >> 4) + 1;                                                      \
        const size_t smem16 = (size_t)L * PITCH * sizeof(float2);                                                    \
        static unsigned long long set_ = 0; const unsigned long long devbit_ = mlb::device_bit();                                                                                    \
        if (!(set_ & devbit_)) {                                                                                                 \
            MLB_CUDA(cudaFuncSetAttribute(mlb::fft16_rows_kernel<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          (int)smem16));                                                             \
            set_ |= devbit_;                                                                                             \
        }                                                                                                            \
        dim3 g16((n_rows + L - 1) / L, batch);                                                                       \
        static unsigned long long set3_ = 0, set4_ = 0;                                                                    \
        if (mlb::g_r16_occ == 3 || ((mlb::g_r16_occ == 2 || mlb::g_r16_occ == 0) && LG == 13)) {                                              \
            constexpr int MB = (LG == 13) ? 2 : 3;                                                                   \
            if (!(set3_ & devbit_)) {                                                                                            \
                MLB_CUDA(cudaFuncSetAttribute(mlb::fft16_rows_kernel<LG, MB>,                                        \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));            \
                set3_ |= devbit_;                                                                                        \
            }                                                                                                        \
            mlb::fft16_rows_kernel<LG, MB><<<g16, T, smem16, st>>>(r);                                               \
        } else if ((mlb::g_r16_occ == 4 || mlb::g_r16_occ == 0) && LG < 13) {                                                                 \
            if (!(set4_ & devbit_)) {                                                                                            \
                MLB_CUDA(cudaFuncSetAttribute(mlb::fft16_rows_kernel<LG, 4>,                                         \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));            \
                set4_ |= devbit_;                                                                                        \
            }                                                                                                        \
            mlb::fft16_rows_kernel<LG, 4><<<g16, T, smem16, st>>>(r);                                                \
        } else {                                                                                                     \
            mlb::fft16_rows_kernel<LG><<<g16, T, smem16, st>>>(r);                                                   \
        }                                                                                                            \
        return mlb::check_launch("mlb_fft_rows(radix 16)");                                                          \
    }
        switch (a.lgN) { MLB_R16_ROWS(8) MLB_R16_ROWS(9) MLB_R16_ROWS(10) MLB_R16_ROWS(11) MLB_R16_ROWS(12) MLB_R16_ROWS(13) }
#undef MLB_R16_ROWS
    }
    // TMA-fed persistent kernel (256..2048 points): needs 16-byte aligned row segments
    {
        bool tma_ok = mlb::g_rows_tma && a.lgN >= 8 && a.lgN <= 11 && (ld_in % 2 == 0);
        for (int b = 0; b < batch; ++b) tma_ok = tma_ok && mlb::aligned16(a.in[b]) && a.in[b] != a.out[b];
        if (tma_ok) {
            static int n_sm = 0;
            if (!n_sm) {
                int dev = 0;
                MLB_CUDA(cudaGetDevice(&dev));
                MLB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
            }
            cudaStream_t st = (cudaStream_t)stream;
#define MLB_ROWS_TMA_RING(LG, KB)                                                                                     \
    {                                                                                                                \
        using Cfg = mlb::RowsTmaCfg<LG, KB>;                                                                         \
        static unsigned long long set_ = 0; const unsigned long long devbit_ = mlb::device_bit();                                                                                    \
        if (!(set_ & devbit_)) {                                                                                                 \
            MLB_CUDA(cudaFuncSetAttribute(mlb::fft_rows_tma_kernel<LG, KB>,                                          \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));             \
            set_ |= devbit_;                                                                                             \
        }                                                                                                            \
        int per_sm = (int)((220 * 1024) / (Cfg::SMEM + 1024));                                                       \
        if (per_sm > 3) per_sm = 3;                                                                                  \
        if (mlb::g_rows_per_sm > 0 && mlb::g_rows_per_sm < per_sm) per_sm = mlb::g_rows_per_sm;                      \
        if (per_sm < 1) per_sm = 1;                                                                                  \
        int grid = n_sm * per_sm;                                                                                    \
        if (grid > n_rows * batch) grid = n_rows * batch;                                                            \
        mlb::fft_rows_tma_kernel<LG, KB><<<grid, mlb::TMA_CONSUMERS + 32, Cfg::SMEM, st>>>(a, batch, sc);            \
        return mlb::check_launch("mlb_fft_rows(tma)");                                                               \
    }
#define MLB_ROWS_TMA(LG)                                                                                              \
    case LG:                                                                                                         \
        if (mlb::g_rows_ring_kb == 128) MLB_ROWS_TMA_RING(LG, 128)                                                   \
        else MLB_ROWS_TMA_RING(LG, 64)
            switch (a.lgN) { MLB_ROWS_TMA(8) MLB_ROWS_TMA(9) MLB_ROWS_TMA(10) MLB_ROWS_TMA(11) }
#undef MLB_ROWS_TMA
#undef MLB_ROWS_TMA_RING
        }
    }
    if (scatter) {
        mlb::set_error("mlb_fft_rows_scatter: operands must be 16-byte aligned and distinct from the outputs");
        return MLB_ERR_UNSUPPORTED;
    }
    MLB_REQUIRE(!transpose_out, "mlb_fft_rows: transposed output needs the TMA-fed kernel (256..2048 points, "
                                "16-byte aligned even-pitch input, out != in); see mlb_fft_rows_can_transpose");
    // Compile-time-sized kernels for 256..8192 points (always 256 threads and max(1, 1024/N) rows per CTA);
    // the runtime-sized kernel covers the small transforms and the tuning knobs.
    const int lgN = a.lgN;
    const bool ct = lgN >= 8 && lgN <= 13 && mlb::g_rows_threads == 256 && mlb::g_rows_points == 1024 &&
                    lanes == ((1024 >> lgN) > 0 ? (1024 >> lgN) : 1);
    cudaStream_t st = (cudaStream_t)stream;
    const int smem_max = 3 * mlb::FFT_MAX_N * 8;
#define MLB_ROWS_LAUNCH(V, LG)                                                                                      \
    do {                                                                                                            \
        static unsigned long long set_ = 0; const unsigned long long devbit_ = mlb::device_bit();                                                                                   \
        if (!(set_ & devbit_)) {                                                                                                \
            MLB_CUDA(cudaFuncSetAttribute(mlb::fft_rows_kernel<V, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          smem_max));                                                               \
            set_ |= devbit_;                                                                                            \
        }                                                                                                           \
        mlb::fft_rows_kernel<V, LG><<<grid, (LG >= 12) ? 1024 : mlb::g_rows_threads, smem, st>>>(a);                                    \
    } while (0)
#define MLB_ROWS_CASE(LG)                                  \
    case LG:                                               \
        if (vec) MLB_ROWS_LAUNCH(2, LG);                   \
        else MLB_ROWS_LAUNCH(1, LG);                       \
        break;
    if (ct) {
        switch (lgN) {
            MLB_ROWS_CASE(8) MLB_ROWS_CASE(9) MLB_ROWS_CASE(10) MLB_ROWS_CASE(11) MLB_ROWS_CASE(12) MLB_ROWS_CASE(13)
        }
    } else {
        if (vec) MLB_ROWS_LAUNCH(2, 0);
        else MLB_ROWS_LAUNCH(1, 0);
    }
#undef MLB_ROWS_CASE
#undef MLB_ROWS_LAUNCH
    return mlb::check_launch("mlb_fft_rows");
}

extern "C" int mlb_fft_rows(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int n_rows,
                            int N, int s1, int s2, const mlb_c64 *tw, int in_roll_r, int in_roll_c, int out_roll,
                            int transpose_out, int batch, void *stream) {
    return fft_rows_impl(h_in, ld_in, h_out, ld_out, n_rows, N, s1, s2, tw, in_roll_r, in_roll_c, out_roll, transpose_out,
                         batch, stream);
}

extern "C" int mlb_fft_rows_scatter(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out_peers, int ld_out,
                                    int n_rows, int N, int s1, int s2, const mlb_c64 *tw, int in_roll_c, int out_roll,
                                    int out_row0, int row_block, int row_stride, int n_rows_total, int world, int batch,
                                    void *stream) {
    MLB_REQUIRE(h_out_peers && world >= 1 && world <= MLB_MAX_PEERS && (world & (world - 1)) == 0,
                "mlb_fft_rows_scatter: world %d must be a power of two <= %d", world, MLB_MAX_PEERS);
    MLB_REQUIRE(batch >= 1 && batch <= 4 && N > 0 && N % world == 0, "mlb_fft_rows_scatter: bad batch %d / N %d", batch, N);
    MLB_REQUIRE(n_rows > 0 && n_rows <= n_rows_total && out_row0 >= 0 && out_row0 < n_rows_total,
                "mlb_fft_rows_scatter: bad row window (%d rows at %d of %d)", n_rows, out_row0, n_rows_total);
    MLB_REQUIRE(row_block >= 1 && (row_block & (row_block - 1)) == 0 && n_rows % row_block == 0 && row_stride >= row_block &&
                    (long long)(n_rows / row_block - 1) * row_stride + row_block <= n_rows_total,
                "mlb_fft_rows_scatter: bad row blocks (%d rows in blocks of %d, %d apart, of %d)", n_rows, row_block,
                row_stride, n_rows_total);
    mlb::RowScatter sc = {};
    sc.world = world; sc.lg_slab = mlb::ilog2(N / world); sc.out_row0 = out_row0; sc.rows_total = n_rows_total;
    sc.lg_block = mlb::ilog2(row_block); sc.row_stride = row_stride;
    for (int p = 0; p < world; ++p)
        for (int f = 0; f < 4; ++f) {
            const mlb_c64 *ptr = h_out_peers[p * batch + (f < batch ? f : 0)];
            MLB_REQUIRE(ptr != nullptr, "mlb_fft_rows_scatter: NULL output (peer %d, field %d)", p, f);
            sc.out[p][f] = reinterpret_cast<float2 *>(const_cast<mlb_c64 *>(ptr));
        }
    return fft_rows_impl(h_in, ld_in, h_out_peers, ld_out, n_rows, N, s1, s2, tw, 0, in_roll_c, out_roll, 0, batch, stream,
                         &sc);
}

namespace mlb {
template <int LG, int CL>
static int launch_r16_cols(const R16ColArgs &r, int groups, int batch, cudaStream_t st) {
    constexpr int PITCH = (1 << LG) + ((1 << LG) >> 4) + 32 / CL, T = CL * ((1 << LG) >> 4);
    const size_t smem16 = (size_t)CL * PITCH * sizeof(float2);
    static unsigned long long set_ = 0, set3_ = 0, set4_ = 0; const unsigned long long devbit_ = mlb::device_bit();
    dim3 gc((r.n_cols + CL - 1) / CL, groups, batch);
    if ((g_r16_occ == 3 || g_r16_occ == 0) && T <= 256) {
        if (!(set3_ & devbit_)) {
            MLB_CUDA(cudaFuncSetAttribute(fft16_cols_kernel<LG, CL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            set3_ |= devbit_;
        }
        fft16_cols_kernel<LG, CL, 3><<<gc, T, smem16, st>>>(r);
    } else if (g_r16_occ == 4 && T <= 256) {
        if (!(set4_ & devbit_)) {
            MLB_CUDA(cudaFuncSetAttribute(fft16_cols_kernel<LG, CL, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            set4_ |= devbit_;
        }
        fft16_cols_kernel<LG, CL, 4><<<gc, T, smem16, st>>>(r);
    } else {
        if (!(set_ & devbit_)) {
            MLB_CUDA(cudaFuncSetAttribute(fft16_cols_kernel<LG, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            set_ |= devbit_;
        }
        fft16_cols_kernel<LG, CL><<<gc, T, smem16, st>>>(r);
    }
    return check_launch("mlb_fft_cols(radix 16)");
}
static int r16_cols_dispatch(int lgSub, const R16ColArgs &r, int groups, int batch, cudaStream_t st) {
    switch (lgSub) {
        case 8: return launch_r16_cols<8, 16>(r, groups, batch, st);
        case 9: return launch_r16_cols<9, 8>(r, groups, batch, st);
        case 10: return launch_r16_cols<10, 4>(r, groups, batch, st);
        case 11: return launch_r16_cols<11, 4>(r, groups, batch, st);
    }
    set_error("radix-16 column pass: unsupported sub-length 2^%d", lgSub);
    return MLB_ERR_ARG;
}
// column strip of the two-pass (>= 4096-point) column transforms: all `fields` fields of a strip stay in L2
// between the passes (multiple of 32 columns = 256-byte row segments)
static int r16_strip_cols(int N, int n_cols, int fields) {
    if (g_cols_strip_mb <= 0) return n_cols;
    long long w = (long long)g_cols_strip_mb * (1 << 20) / ((long long)fields * N * 8);
    w = w / 32 * 32;
    if (w < 32) w = 32;
    return w < n_cols ? (int)w : n_cols;
}

template <int LG, int CL>
static int launch_r16_cols_power(const R16PowerArgs &r, int groups, cudaStream_t st) {
    constexpr int PITCH = (1 << LG) + ((1 << LG) >> 4) + 32 / CL, T = CL * ((1 << LG) >> 4);
    const size_t smem16 = (size_t)CL * PITCH * sizeof(float2) + (size_t)16 * T * (sizeof(float2) + sizeof(float));
    static unsigned long long set_ = 0; const unsigned long long devbit_ = mlb::device_bit();
    if (!(set_ & devbit_)) {
        MLB_CUDA(cudaFuncSetAttribute(fft16_cols_power_kernel<LG, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
        set_ |= devbit_;
    }
    dim3 gc((r.n_cols + CL - 1) / CL, groups);
    fft16_cols_power_kernel<LG, CL><<<gc, T, smem16, st>>>(r);
    return check_launch("mlb_fft_cols_power(radix 16)");
}
static int r16_cols_power_dispatch(int lgSub, const R16PowerArgs &r, int groups, cudaStream_t st) {
    switch (lgSub) {
        case 8: return launch_r16_cols_power<8, 16>(r, groups, st);
        case 9: return launch_r16_cols_power<9, 8>(r, groups, st);
        case 10: return launch_r16_cols_power<10, 4>(r, groups, st);
        case 11: return launch_r16_cols_power<11, 2>(r, groups, st);
    }
    set_error("radix-16 column+power pass: unsupported sub-length 2^%d", lgSub);
    return MLB_ERR_ARG;
}
}  // namespace mlb

extern "C" int mlb_fft_cols(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int N,
                            int n_cols, const mlb_c64 *tw, int out_roll, int batch, void *stream) {
    mlb::FftArgs a;
    if (int rc = mlb::fill_args(a, h_in, h_out, batch, "mlb_fft_cols")) return rc;
    int radix_probe[16];
    MLB_REQUIRE(N >= 2 && N <= mlb::FFT_MAX_N && mlb::factor_235(N, radix_probe) > 0,
                "mlb_fft_cols: length %d must be of the form 2^a 3^b 5^c and <= %d", N, mlb::FFT_MAX_N);
    MLB_REQUIRE(tw && n_cols > 0 && ld_in >= n_cols && ld_out >= n_cols, "mlb_fft_cols: bad sizes");
    MLB_REQUIRE(out_roll >= 0 && out_roll < N, "mlb_fft_cols: roll out of range");
    a.tw = reinterpret_cast<const float2 *>(tw);
    if (!mlb::is_pow2(N)) {
        // big-radix engine: direct while >= 16 columns per CTA fit, else N = A x B in two passes (the first in place
        // on the INPUT buffer, which is then scratch; in-place calls fall back to the direct pass)
        cudaStream_t st = (cudaStream_t)stream;
        bool inplace = false;
        for (int b = 0; b < batch; ++b) inplace = inplace || (a.in[b] == a.out[b]);
        int A = 1;
        if (!inplace && mlb::mix2_col_lanes(N, n_cols) < 16 && n_cols >= 8) {
            static const int cand[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2};
            for (int c : cand)
                if (N % c == 0 && N / c >= 2) { A = c; break; }
        }
        const int B = N / A;
        mlb::Mix2Args m2;
        MLB_REQUIRE(mlb::mix2_plan(m2, B) > 0, "mlb_fft_cols: cannot factor %d", B);
        for (int b = 0; b < 4; ++b) { m2.in[b] = a.in[b]; m2.out[b] = a.out[b]; }
        m2.tw = a.tw; m2.ld_in = ld_in; m2.ld_out = ld_out; m2.other = n_cols; m2.tw_mul = A; m2.Ntot = N;
        m2.in_roll_r = m2.in_roll_c = m2.out_roll = 0; m2.s1 = m2.s2 = 1; m2.roll = out_roll;
        if (A > 1) {
            mlb::Mix2FirstArgs f;
            for (int b = 0; b < 4; ++b) { f.in[b] = a.in[b]; f.out[b] = const_cast<float2 *>(a.in[b]); }
            f.tw = a.tw; f.ld = ld_in; f.n_cols = n_cols; f.B = B;
            if (int rc = mlb::mix2_first(A, f, batch, st)) return rc;
            m2.in_gs = B; m2.in_rs = 1; m2.out_gs = 1; m2.out_rs = A;
        } else {
            m2.in_gs = 0; m2.in_rs = 1; m2.out_gs = 0; m2.out_rs = 1;
        }
        if (int rc = mlb::mix2_set_smem()) return rc;
        const int mp = mlb::mix2_max_per(m2);
        int rl = 1;
        while (rl < 32 && rl * 2 * mp <= 256) rl *= 2;
        if (mlb::g_mixed_reg && rl * mp <= 256 && rl >= 8 && (mlb::g_mixed_reg == 2 || rl * mp >= 205)) {
            m2.lanes = rl; m2.lg_lanes = mlb::ilog2(rl);   // one butterfly per thread and stage: register variant
            dim3 gr((n_cols + rl - 1) / rl, A, batch);
            const size_t sm2 = mlb::mix2_smem(B, rl) / 2;
            if (mlb::Mix2Kernel ct = (sm2 <= 48 * 1024) ? mlb::mix2_ct_find(mlb::mix2_ct_cols, m2, mlb::MIX2_CT_COL_LANES, mlb::g_mixed_occ) : nullptr) {
                ct<<<gr, 256, sm2, st>>>(m2);
                return mlb::check_launch("mlb_fft_cols(mixed radix, compiled plan)");
            }
            if (mlb::g_mixed_occ == 2) mlb::mix2_reg_kernel<true, 2><<<gr, 256, sm2, st>>>(m2);
            else if (mlb::g_mixed_occ == 3) mlb::mix2_reg_kernel<true, 3><<<gr, 256, sm2, st>>>(m2);
            else mlb::mix2_reg_kernel<true, 4><<<gr, 256, sm2, st>>>(m2);
            return mlb::check_launch("mlb_fft_cols(mixed radix, registers)");
        }
        const int lanes = mlb::mix2_col_lanes(B, n_cols);
        m2.lanes = lanes; m2.lg_lanes = mlb::ilog2(lanes);
        dim3 grid((n_cols + lanes - 1) / lanes, A, batch);
        mlb::mix2_cols_kernel<<<grid, 256, mlb::mix2_smem(B, lanes), st>>>(m2);
        return mlb::check_launch("mlb_fft_cols(mixed radix)");
    }
    a.ld_in = ld_in; a.ld_out = ld_out; a.lgN = mlb::ilog2(N); a.other = n_cols;
    a.in_roll_r = 0; a.in_roll_c = 0; a.out_roll = out_roll; a.s1 = a.s2 = 1; a.evict_first = 0;
    if (mlb::g_cols_engine == 1 && a.lgN >= mlb::g_r16_min_lg && a.lgN <= 13) {
        // radix-16 register kernels: direct up to 2048 points; 4096 / 8192 = 16 x (256 / 512) in two passes run
        // strip by strip, the first of them writing into the first strip of the INPUT buffer (which is therefore
        // scratch, as with the radix-4 four-step)
        cudaStream_t st = (cudaStream_t)stream;
        mlb::R16ColArgs r;
        for (int b = 0; b < 4; ++b) { r.in[b] = a.in[b]; r.out[b] = a.out[b]; }
        r.ld_in = ld_in; r.ld_out = ld_out; r.n_cols = n_cols; r.lgNtot = a.lgN; r.roll = out_roll;
        r.tw = a.tw;
        if (a.lgN < 12) {
            r.in_gs = 0; r.in_rs = 1; r.out_gs = 0; r.out_rs = 1;
            return mlb::r16_cols_dispatch(a.lgN, r, 1, batch, st);
        }
        for (int b = 0; b < batch; ++b)
            MLB_REQUIRE(a.in[b] != a.out[b], "mlb_fft_cols: N >= 4096 needs out != in (the input is used as scratch)");
        const int lgSub = a.lgN - 4;
        r.in_gs = 1 << lgSub; r.in_rs = 1; r.out_gs = 1; r.out_rs = 16;
        const int strip = mlb::r16_strip_cols(N, n_cols, batch);
        for (int c0 = 0; c0 < n_cols; c0 += strip) {
            const int w = (n_cols - c0 < strip) ? n_cols - c0 : strip;
            mlb::R16FirstArgs f;
            for (int b = 0; b < 4; ++b) { f.in[b] = a.in[b] + c0; f.out[b] = const_cast<float2 *>(a.in[b]); r.out[b] = a.out[b] + c0; }
            f.tw = a.tw; f.ld = ld_in; f.n_cols = w; f.B = 1 << lgSub;
            dim3 gf((w + 31) / 32, (f.B + 7) / 8, batch);
            mlb::fft16_cols_first_kernel<<<gf, 256, 0, st>>>(f);
            if (int rc = mlb::check_launch("mlb_fft_cols(radix-16 first pass)")) return rc;
            r.n_cols = w;
            if (int rc = mlb::r16_cols_dispatch(lgSub, r, 16, batch, st)) return rc;
        }
        return MLB_OK;
    }
    if (a.lgN >= 12) {
        for (int b = 0; b < batch; ++b)
            MLB_REQUIRE(a.in[b] != a.out[b], "mlb_fft_cols: N >= 4096 needs out != in (the input is used as scratch)");
        // four-step: N = A*B with A = 64; pass 1 works in place on the INPUT buffer (it is overwritten)
        const int lgA = 6, lgB = a.lgN - lgA, Asz = 1 << lgA, Bsz = 1 << lgB;
        mlb::FftSubArgs s1, s2;
        for (int b = 0; b < 4; ++b) {
            s1.in[b] = a.in[b]; s1.out[b] = const_cast<float2 *>(a.in[b]);
            s2.in[b] = a.in[b]; s2.out[b] = a.out[b];
        }
        s1.tw = s2.tw = a.tw;
        s1.ld_in = s1.ld_out = ld_in; s2.ld_in = ld_in; s2.ld_out = ld_out;
        s1.lgNtot = s2.lgNtot = a.lgN; s1.n_cols = s2.n_cols = n_cols;
        s1.in_gs = 1; s1.in_rs = Bsz; s1.out_gs = 1; s1.out_rs = Bsz; s1.roll = 0; s1.twiddle = 1;
        s2.in_gs = Bsz; s2.in_rs = 1; s2.out_gs = 1; s2.out_rs = Asz; s2.roll = out_roll; s2.twiddle = 0;
        cudaStream_t st = (cudaStream_t)stream;
        dim3 g1((n_cols + 15) / 16, Bsz, batch), g2((n_cols + 15) / 16, Asz, batch);
        mlb::fft_cols_sub_kernel<6, 16><<<g1, 256, 0, st>>>(s1);
        if (int rc = mlb::check_launch("mlb_fft_cols(four-step pass 1)")) return rc;
        if (lgB == 6) mlb::fft_cols_sub_kernel<6, 16><<<g2, 256, 0, st>>>(s2);
        else mlb::fft_cols_sub_kernel<7, 16><<<g2, 256, 0, st>>>(s2);
        return mlb::check_launch("mlb_fft_cols(four-step pass 2)");
    }
    // as many adjacent columns as fit 64 KB (so 3 CTAs share an SM), at most 16 (128-byte row segments)
    int lanes = 1;
    while (lanes < 16 && 2 * (size_t)(lanes * 2) * (N + 4) * sizeof(float2) <= 66 * 1024 && lanes * 2 <= n_cols) lanes *= 2;
    if (mlb::g_cols_half && lanes >= 2) lanes /= 2;      // tuning: narrower column tiles, twice the CTAs per SM
    a.lanes = lanes;
    const size_t smem = (2 * (size_t)lanes * (N + 4) + N) * sizeof(float2);
    cudaStream_t st = (cudaStream_t)stream;
    const int smem_max = (3 * mlb::FFT_MAX_N + 8) * 8;
#define MLB_COLS_LAUNCH(LG, CL)                                                                                      \
    do {                                                                                                             \
        static unsigned long long set_ = 0; const unsigned long long devbit_ = mlb::device_bit();                                                                                    \
        if (!(set_ & devbit_)) {                                                                                                 \
            MLB_CUDA(cudaFuncSetAttribute(mlb::fft_cols_kernel<LG, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          smem_max));                                                                \
            set_ |= devbit_;                                                                                             \
        }                                                                                                            \
        dim3 grid((n_cols + lanes - 1) / lanes, batch);                                                              \
        mlb::fft_cols_kernel<LG, CL><<<grid, (LG >= 12) ? 1024 : 256, smem, st>>>(a);                                                    \
    } while (0)
    // compile-time-sized kernels when the lane count is the canonical one for that length
    if (a.lgN == 8 && lanes == 16) MLB_COLS_LAUNCH(8, 16);
    else if (a.lgN == 9 && lanes == 8) MLB_COLS_LAUNCH(9, 8);
    else if (a.lgN == 10 && lanes == 4) MLB_COLS_LAUNCH(10, 4);
    else if (a.lgN == 10 && lanes == 2) MLB_COLS_LAUNCH(10, 2);
    else if (a.lgN == 9 && lanes == 4) MLB_COLS_LAUNCH(9, 4);
    else if (a.lgN == 11 && lanes == 1) MLB_COLS_LAUNCH(11, 1);
    else if (a.lgN == 11 && lanes == 2) MLB_COLS_LAUNCH(11, 2);
    else if (a.lgN == 12 && lanes == 1) MLB_COLS_LAUNCH(12, 1);
    else if (a.lgN == 13 && lanes == 1) MLB_COLS_LAUNCH(13, 1);
    else MLB_COLS_LAUNCH(0, 0);
#undef MLB_COLS_LAUNCH
    return mlb::check_launch("mlb_fft_cols");
}

// column tile width of the fused column+power pass for a length-N transform (0: not supported)
static int cols_power_tile(int N) {
    if (!mlb::is_pow2(N) || N < 256 || N > 2048) return 0;
    const bool wide = mlb::g_cols_power_wide < 0 ? (N >= 1024) : (mlb::g_cols_power_wide != 0);
    return (wide ? 4096 : 2048) / N;
}
// radix-16 engine: sub-transform length (N itself up to 2048, N/16 above), its column tile and group count
static bool r16_power_shape(int N, int *lgSub, int *cl, int *groups) {
    if (!mlb::is_pow2(N) || N < 256 || N > 8192) return false;
    const int lg = mlb::ilog2(N);
    *lgSub = lg >= 12 ? lg - 4 : lg;
    *groups = lg >= 12 ? 16 : 1;
    *cl = 4096 >> *lgSub;                       // 256 threads per CTA: 16, 8, 4, 2 columns
    return true;
}

extern "C" int mlb_fft_cols_power_blocks(int N, int n_cols) {
    if (n_cols <= 0) return 0;
    if (mlb::g_cols_engine == 1 && N >= (1 << mlb::g_r16_min_lg)) {
        int lgSub, cl, groups;
        if (!r16_power_shape(N, &lgSub, &cl, &groups)) return 0;
        return groups * ((n_cols + cl - 1) / cl);
    }
    const int cl = cols_power_tile(N);
    return cl ? (n_cols + cl - 1) / cl : 0;
}

static int fft_cols_power_impl(const mlb_c64 *const *h_in, int ld_in, int N, int n_cols, const mlb_c64 *tw,
                               int out_roll, const double *ux, const double *uy, double amp_scale,
                               double wavelength, double n_glass, double Z0, float *P, int ldp, int accumulate,
                               double *block_sums, mlb_c64 *const *h_Fhat, int ldf, void *stream,
                               double *total, double total_scale, unsigned int *done, bool *total_done) {
    if (total_done) *total_done = false;
    MLB_REQUIRE(h_in && tw && ux && uy && P, "mlb_fft_cols_power: NULL pointer");
    if (mlb::g_cols_engine == 1 && !h_Fhat && N >= (1 << mlb::g_r16_min_lg)) {
        int lgSub, cl16, groups;
        MLB_REQUIRE(r16_power_shape(N, &lgSub, &cl16, &groups), "mlb_fft_cols_power: length %d must be a power of two in 256..8192", N);
        MLB_REQUIRE(n_cols > 0 && ld_in >= n_cols && ldp >= n_cols && out_roll >= 0 && out_roll < N, "mlb_fft_cols_power: bad sizes");
        MLB_REQUIRE(wavelength > 0 && n_glass > 0 && Z0 > 0, "mlb_fft_cols_power: bad physical constants");
        cudaStream_t st16 = (cudaStream_t)stream;
        mlb::R16PowerArgs r;
        for (int f = 0; f < 4; ++f) {
            MLB_REQUIRE(h_in[f] != nullptr, "mlb_fft_cols_power: field %d is NULL", f);
            r.in[f] = reinterpret_cast<const float2 *>(h_in[f]);
        }
        r.tw = reinterpret_cast<const float2 *>(tw);
        r.ux = ux; r.uy = uy; r.P = P; r.block_sums = block_sums;
        const double pi16 = 3.14159265358979323846, Zd16 = Z0 / n_glass, k16 = 2 * pi16 * n_glass / wavelength;
        r.scale = k16 * k16 / (32 * pi16 * pi16 * Zd16) * amp_scale * amp_scale * 2.0;
        r.Z = (float)Zd16;
        r.ld_in = ld_in; r.ldp = ldp; r.n_cols = n_cols; r.lgNtot = mlb::ilog2(N); r.accumulate = accumulate ? 1 : 0;
        r.roll = out_roll;
        const int xblocks = (n_cols + cl16 - 1) / cl16;
        r.bs_stride = xblocks; r.bs_off = 0;
        r.total = nullptr; r.total_scale = 0.0; r.done = nullptr; r.n_sums = 0;
        if (groups == 1) {
            r.in_gs = 0; r.in_rs = 1; r.out_gs = 0; r.out_rs = 1;
            if (total && done && block_sums) {                      // the last CTA finishes total_P (256-thread CTAs)
                r.total = total; r.total_scale = total_scale; r.done = done; r.n_sums = xblocks;
                if (total_done) *total_done = true;
            }
            return mlb::r16_cols_power_dispatch(lgSub, r, 1, st16);
        }
        // 16 x B decomposition, strip by strip: the first pass of every strip writes into the first strip of the
        // inputs (scratch), the fused second pass reads it back from L2
        r.in_gs = 1 << lgSub; r.in_rs = 1; r.out_gs = 1; r.out_rs = 16;
        const int strip = mlb::r16_strip_cols(N, n_cols, 4);
        for (int c0 = 0; c0 < n_cols; c0 += strip) {
            const int w = (n_cols - c0 < strip) ? n_cols - c0 : strip;
            mlb::R16FirstArgs f1;
            for (int f = 0; f < 4; ++f) { f1.in[f] = r.in[f] + c0; f1.out[f] = const_cast<float2 *>(r.in[f]); }
            f1.tw = r.tw; f1.ld = ld_in; f1.n_cols = w; f1.B = 1 << lgSub;
            dim3 gf((w + 31) / 32, (f1.B + 7) / 8, 4);
            mlb::fft16_cols_first_kernel<<<gf, 256, 0, st16>>>(f1);
            if (int rc = mlb::check_launch("mlb_fft_cols_power(radix-16 first pass)")) return rc;
            mlb::R16PowerArgs rs = r;
            rs.n_cols = w; rs.uy = uy + c0; rs.P = P + c0; rs.bs_off = c0 / cl16;
            if (int rc = mlb::r16_cols_power_dispatch(lgSub, rs, 16, st16)) return rc;
        }
        return MLB_OK;
    }
    const int cl = cols_power_tile(N);
    MLB_REQUIRE(cl > 0, "mlb_fft_cols_power: length %d must be a power of two in 256..2048", N);
    MLB_REQUIRE(n_cols > 0 && ld_in >= n_cols && ldp >= n_cols && out_roll >= 0 && out_roll < N,
                "mlb_fft_cols_power: bad sizes");
    MLB_REQUIRE(wavelength > 0 && n_glass > 0 && Z0 > 0, "mlb_fft_cols_power: bad physical constants");
    mlb::ColsPowerArgs a;
    for (int f = 0; f < 4; ++f) {
        MLB_REQUIRE(h_in[f] != nullptr, "mlb_fft_cols_power: field %d is NULL", f);
        a.in[f] = reinterpret_cast<const float2 *>(h_in[f]);
        a.fhat[f] = nullptr;
        if (h_Fhat) {
            MLB_REQUIRE(h_Fhat[f] != nullptr && ldf >= n_cols, "mlb_fft_cols_power: bad Fhat operand %d", f);
            a.fhat[f] = reinterpret_cast<float2 *>(h_Fhat[f]);
        }
    }
    a.store_fhat = h_Fhat ? 1 : 0;
    a.tw = reinterpret_cast<const float2 *>(tw);
    a.ux = ux; a.uy = uy; a.P = P; a.block_sums = block_sums;
    const double pi = 3.14159265358979323846;
    const double Zd = Z0 / n_glass;                                            // nearfield_farfield.py:183
    const double k = 2 * pi * n_glass / wavelength;
    a.scale = k * k / (32 * pi * pi * Zd) * amp_scale * amp_scale * 2.0;       // :184, :189
    a.Z = (float)Zd;
    a.ld_in = ld_in; a.ldf = ldf; a.ldp = ldp; a.n_cols = n_cols; a.out_roll = out_roll; a.accumulate = accumulate ? 1 : 0;
    const int lgN = mlb::ilog2(N);
    const int grid = (n_cols + cl - 1) / cl;
    const size_t smem = (2 * (size_t)cl * (N + 4) + N) * sizeof(float2) + (size_t)N * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define MLB_CP_LAUNCH(LG, CL, T)                                                                                      \
    do {                                                                                                              \
        static unsigned long long set_ = 0; const unsigned long long devbit_ = mlb::device_bit();                                                                                     \
        if (!(set_ & devbit_)) {                                                                                                  \
            MLB_CUDA(cudaFuncSetAttribute(mlb::fft_cols_power_kernel<LG, CL, T>,                                      \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                   \
            set_ |= devbit_;                                                                                              \
        }                                                                                                             \
        mlb::fft_cols_power_kernel<LG, CL, T><<<grid, T, smem, st>>>(a);                                              \
    } while (0)
    if (cl * N == 2048) {                                // 2048-point tiles, 512 threads, 4 points per thread and field
        switch (lgN) {
            case 8: MLB_CP_LAUNCH(8, 8, 512); break;
            case 9: MLB_CP_LAUNCH(9, 4, 512); break;
            case 10: MLB_CP_LAUNCH(10, 2, 512); break;
            default: MLB_CP_LAUNCH(11, 1, 512); break;
        }
    } else {                                             // 4096-point tiles, 1024 threads
        switch (lgN) {
            case 8: MLB_CP_LAUNCH(8, 16, 1024); break;
            case 9: MLB_CP_LAUNCH(9, 8, 1024); break;
            case 10: MLB_CP_LAUNCH(10, 4, 1024); break;
            default: MLB_CP_LAUNCH(11, 2, 1024); break;
        }
    }
#undef MLB_CP_LAUNCH
    return mlb::check_launch("mlb_fft_cols_power");
}

extern "C" int mlb_fft_cols_power(const mlb_c64 *const *h_in, int ld_in, int N, int n_cols, const mlb_c64 *tw,
                                  int out_roll, const double *ux, const double *uy, double amp_scale,
                                  double wavelength, double n_glass, double Z0, float *P, int ldp, int accumulate,
                                  double *block_sums, mlb_c64 *const *h_Fhat, int ldf, void *stream) {
    return fft_cols_power_impl(h_in, ld_in, N, n_cols, tw, out_roll, ux, uy, amp_scale, wavelength, n_glass, Z0, P, ldp,
                               accumulate, block_sums, h_Fhat, ldf, stream, nullptr, 0.0, nullptr, nullptr);
}

extern "C" int mlb_fft_cols_power_total(const mlb_c64 *const *h_in, int ld_in, int N, int n_cols, const mlb_c64 *tw,
                                        int out_roll, const double *ux, const double *uy, double amp_scale,
                                        double wavelength, double n_glass, double Z0, float *P, int ldp, int accumulate,
                                        double *block_sums, double *total, double total_scale, void *done_counter,
                                        void *stream) {
    MLB_REQUIRE(block_sums && total && done_counter, "mlb_fft_cols_power_total: NULL block_sums / total / counter");
    bool fused = false;
    if (int rc = fft_cols_power_impl(h_in, ld_in, N, n_cols, tw, out_roll, ux, uy, amp_scale, wavelength, n_glass, Z0, P,
                                     ldp, accumulate, block_sums, nullptr, 0, stream, total, total_scale,
                                     reinterpret_cast<unsigned int *>(done_counter), &fused))
        return rc;
    if (fused) return MLB_OK;
    return mlb_sum_f64(block_sums, mlb_fft_cols_power_blocks(N, n_cols), total_scale, total, stream);
}
