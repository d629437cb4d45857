// Shared-memory FFT passes for power-of-two FFT-bin far-field grids.
//
// The reference's own algorithm is fft2(fftshift(J)) (nearfield_farfield.py:18-20).  For grids that
// are (a stride of) the FFT bins, the aperture sum is a 2-D DFT of the (folded) aperture, done here
// as two streaming passes -- rows, then columns -- each a Stockham autosort radix-4 (+ one radix-2)
// FFT entirely inside shared memory: one read and one write of the data per pass.
//   * the row pass folds the aperture while loading (sum of the s1*s2 aliased samples per point,
//     see fold.cu), so for a strided grid the full aperture is read from HBM exactly once and the
//     folded aperture never goes to memory: this kernel is the HBM-bound hot kernel of NF->FF;
//   * fftshift of input and output is index arithmetic ("rolls") at load/store time;
//   * twiddles come from a table built with float64 phases.
// All sizes are powers of two, so index math is shifts and masks.
#include "common.cuh"

namespace mlb {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// One Stockham stage of radix R over 2^lgL independent transforms of length 2^lgN in shared memory.
// Element n of transform `lane` lives at  COLS ? (n << lgL) + lane : (lane << lgN) + n.
template <int R, bool COLS>
__device__ __forceinline__ void stockham_stage(const float2 *__restrict__ x, float2 *__restrict__ y, int lgN, int lgNs,
                                               int lgL, const float2 *__restrict__ tw) {
    constexpr int lgR = (R == 4) ? 2 : 1;
    const int lgPer = lgN - lgR;
    const int per = 1 << lgPer;
    const int total = per << lgL;
    const int Ns = 1 << lgNs;
    const int lgTstep = lgN - lgNs - lgR;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int j, lane;
        if (COLS) { j = idx >> lgL; lane = idx & ((1 << lgL) - 1); }
        else { lane = idx >> lgPer; j = idx & (per - 1); }
        const int k = j & (Ns - 1);
        const int base_out = ((j - k) << lgR) + k;             // expand(j, Ns, R)
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = j + r * per;
            v[r] = x[COLS ? (n << lgL) + lane : (lane << lgN) + n];
            if (r > 0 && lgNs > 0) v[r] = cmulf(v[r], __ldg(tw + ((r * k) << lgTstep)));
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
        for (int r = 0; r < R; ++r) {
            const int n = base_out + (r << lgNs);
            y[COLS ? (n << lgL) + lane : (lane << lgN) + n] = v[r];
        }
    }
}

// runs all stages; returns the buffer (0 or 1) that holds the result
template <bool COLS>
__device__ __forceinline__ int fft_in_smem(float2 *buf0, float2 *buf1, int lgN, int lgL, const float2 *tw) {
    int cur = 0, lgNs = 0;
    while (lgNs + 2 <= lgN) {
        __syncthreads();
        if (cur == 0) stockham_stage<4, COLS>(buf0, buf1, lgN, lgNs, lgL, tw);
        else stockham_stage<4, COLS>(buf1, buf0, lgN, lgNs, lgL, tw);
        cur ^= 1;
        lgNs += 2;
    }
    if (lgNs < lgN) {
        __syncthreads();
        if (cur == 0) stockham_stage<2, COLS>(buf0, buf1, lgN, lgNs, lgL, tw);
        else stockham_stage<2, COLS>(buf1, buf0, lgN, lgNs, lgL, tw);
        cur ^= 1;
    }
    __syncthreads();
    return cur;
}

struct FftArgs {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;
    int ld_in, ld_out, lgN, other, lgL, in_roll_r, in_roll_c, out_roll, s1, s2;
};

// rows: 2^lgL consecutive rows per CTA, transform along the contiguous axis; the loader sums the
// s1 x s2 aliased copies (aperture fold) and applies the input fftshift.
template <int VEC>
__global__ void __launch_bounds__(256) fft_rows_kernel(const FftArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int N = 1 << a.lgN, L = 1 << a.lgL;
    float2 *buf0 = fsm, *buf1 = fsm + ((size_t)L << a.lgN);
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    const int row0 = blockIdx.x << a.lgL;
    const int total = (L << a.lgN) / VEC;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int e = idx * VEC;
        const int lane = e >> a.lgN, n = e & (N - 1);
        const int r = row0 + lane;
        float acc[2 * VEC];
#pragma unroll
        for (int v = 0; v < 2 * VEC; ++v) acc[v] = 0.f;
        if (r < a.other) {
            int rs = r - a.in_roll_r; if (rs < 0) rs += a.other;
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
        if (VEC == 2) *reinterpret_cast<float4 *>(buf0 + e) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else buf0[e] = make_float2(acc[0], acc[1]);
    }
    const int cur = fft_in_smem<false>(buf0, buf1, a.lgN, a.lgL, a.tw);
    const float2 *res = cur ? buf1 : buf0;
    const int tot = L << a.lgN;
    for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        const int lane = idx >> a.lgN, n = idx & (N - 1);       // n = position in the OUTPUT row
        const int r = row0 + lane;
        if (r < a.other) {
            const int q = (n - a.out_roll) & (N - 1);           // out[(q + roll) % N] = X[q]
            out[(size_t)r * a.ld_out + n] = res[(lane << a.lgN) + q];
        }
    }
}

// columns: 2^lgL adjacent columns per CTA, transform along the strided axis
__global__ void __launch_bounds__(256) fft_cols_kernel(const FftArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    const int N = 1 << a.lgN, L = 1 << a.lgL;
    float2 *buf0 = fsm, *buf1 = fsm + ((size_t)L << a.lgN);
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    const int c0 = blockIdx.x << a.lgL;
    const int total = L << a.lgN;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int n = idx >> a.lgL, lane = idx & (L - 1);
        const int c = c0 + lane;
        buf0[idx] = (c < a.other) ? in[(size_t)n * a.ld_in + c] : make_float2(0.f, 0.f);
    }
    const int cur = fft_in_smem<true>(buf0, buf1, a.lgN, a.lgL, a.tw);
    const float2 *res = cur ? buf1 : buf0;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int n = idx >> a.lgL, lane = idx & (L - 1);       // n = OUTPUT row
        const int c = c0 + lane;
        if (c < a.other) {
            const int q = (n - a.out_roll) & (N - 1);
            out[(size_t)n * a.ld_out + c] = res[(q << a.lgL) + lane];
        }
    }
}

__global__ void fft_twiddle_kernel(int N, float2 *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    double s, c;
    sincospi(-2.0 * (double)t / (double)N, &s, &c);
    out[t] = make_float2((float)c, (float)s);
}

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
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

extern "C" int mlb_fft_max_length(void) { return mlb::FFT_MAX_N; }

extern "C" int mlb_fft_rows(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int n_rows,
                            int N, int s1, int s2, const mlb_c64 *tw, int in_roll_r, int in_roll_c, int out_roll,
                            int batch, void *stream) {
    mlb::FftArgs a;
    if (int rc = mlb::fill_args(a, h_in, h_out, batch, "mlb_fft_rows")) return rc;
    MLB_REQUIRE(mlb::is_pow2(N) && N >= 2 && N <= mlb::FFT_MAX_N, "mlb_fft_rows: length %d must be a power of two <= %d",
                N, mlb::FFT_MAX_N);
    MLB_REQUIRE(s1 >= 1 && s2 >= 1, "mlb_fft_rows: fold factors must be >= 1");
    MLB_REQUIRE(tw && n_rows > 0 && ld_in >= N * s2 && ld_out >= N, "mlb_fft_rows: bad sizes");
    MLB_REQUIRE(in_roll_r >= 0 && in_roll_r < n_rows && in_roll_c >= 0 && in_roll_c < N && out_roll >= 0 && out_roll < N,
                "mlb_fft_rows: rolls out of range");
    for (int b = 0; b < batch; ++b)
        MLB_REQUIRE((in_roll_r == 0 && s1 == 1 && s2 == 1) || a.in[b] != a.out[b],
                    "mlb_fft_rows: in-place needs in_roll_r == 0 and no fold");
    a.tw = reinterpret_cast<const float2 *>(tw);
    a.ld_in = ld_in; a.ld_out = ld_out; a.lgN = mlb::ilog2(N); a.other = n_rows;
    a.in_roll_r = in_roll_r; a.in_roll_c = in_roll_c; a.out_roll = out_roll; a.s1 = s1; a.s2 = s2;
    int lanes = 1;
    while (lanes * 2 * N <= 1024 && lanes * 2 <= n_rows) lanes *= 2;      // small transforms: several rows per CTA
    a.lgL = mlb::ilog2(lanes);
    const size_t smem = 2 * (size_t)lanes * N * sizeof(float2);
    bool vec = (N >= 2) && (in_roll_c % 2 == 0) && (ld_in % 2 == 0);
    for (int b = 0; b < batch; ++b) vec = vec && mlb::aligned16(a.in[b]);
    dim3 grid((n_rows + lanes - 1) / lanes, batch);
    static bool attr_set = false;            // once per process: keeps launches capturable in CUDA graphs
    if (!attr_set) {
        MLB_CUDA(cudaFuncSetAttribute(mlb::fft_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * mlb::FFT_MAX_N * 8));
        MLB_CUDA(cudaFuncSetAttribute(mlb::fft_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * mlb::FFT_MAX_N * 8));
        attr_set = true;
    }
    if (vec) mlb::fft_rows_kernel<2><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
    else mlb::fft_rows_kernel<1><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
    return mlb::check_launch("mlb_fft_rows");
}

extern "C" int mlb_fft_cols(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int N,
                            int n_cols, const mlb_c64 *tw, int out_roll, int batch, void *stream) {
    mlb::FftArgs a;
    if (int rc = mlb::fill_args(a, h_in, h_out, batch, "mlb_fft_cols")) return rc;
    MLB_REQUIRE(mlb::is_pow2(N) && N >= 2 && N <= mlb::FFT_MAX_N, "mlb_fft_cols: length %d must be a power of two <= %d",
                N, mlb::FFT_MAX_N);
    MLB_REQUIRE(tw && n_cols > 0 && ld_in >= n_cols && ld_out >= n_cols, "mlb_fft_cols: bad sizes");
    MLB_REQUIRE(out_roll >= 0 && out_roll < N, "mlb_fft_cols: roll out of range");
    a.tw = reinterpret_cast<const float2 *>(tw);
    a.ld_in = ld_in; a.ld_out = ld_out; a.lgN = mlb::ilog2(N); a.other = n_cols;
    a.in_roll_r = 0; a.in_roll_c = 0; a.out_roll = out_roll; a.s1 = a.s2 = 1;
    // as many adjacent columns as fit 64 KB (so 3 CTAs share an SM), at most 16 (128-byte row segments)
    int lanes = 1;
    while (lanes < 16 && 2 * (size_t)(lanes * 2) * N * sizeof(float2) <= 64 * 1024 && lanes * 2 <= n_cols) lanes *= 2;
    a.lgL = mlb::ilog2(lanes);
    const size_t smem = 2 * (size_t)lanes * N * sizeof(float2);
    static bool attr_set = false;
    if (!attr_set) {
        MLB_CUDA(cudaFuncSetAttribute(mlb::fft_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * mlb::FFT_MAX_N * 8));
        attr_set = true;
    }
    dim3 grid((n_cols + lanes - 1) / lanes, batch);
    mlb::fft_cols_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
    return mlb::check_launch("mlb_fft_cols");
}
