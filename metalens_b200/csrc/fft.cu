// Shared-memory FFT passes for power-of-two FFT-bin far-field grids.
//
// The reference's own algorithm is fft2(fftshift(J)) (nearfield_farfield.py:18-20).  For grids that
// are (a stride of) the FFT bins, the aperture sum is a 2-D DFT of the (folded) aperture, and the
// DFT is done here as two streaming passes -- rows, then columns -- each a Stockham autosort
// radix-4 (+ one radix-2) FFT entirely inside shared memory, one read and one write of the data
// per pass.  Both passes are memory-bound (HBM for the big all-bins case, L2 for folded apertures);
// twiddles come from a float64-accurate table.  fftshift of input and output is index arithmetic
// ("rolls") at load/store time, so no extra pass is spent on it.
#include "common.cuh"

namespace mlb {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// One Stockham stage of radix R over `lanes` independent transforms held in shared memory.
// Element n of transform `lane` lives at  COLS ? n*lanes + lane : lane*N + n.
template <int R, bool COLS>
__device__ __forceinline__ void stockham_stage(const float2 *__restrict__ x, float2 *__restrict__ y, int N, int Ns,
                                               int lanes, const float2 *__restrict__ tw) {
    const int per = N / R;
    const int total = per * lanes;
    const int tstep = N / (Ns * R);
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int j, lane;
        if (COLS) { j = idx / lanes; lane = idx - j * lanes; }
        else { lane = idx / per; j = idx - lane * per; }
        const int k = j % Ns;
        const int base_out = (j - k) * R + k;                  // expand(j, Ns, R)
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = j + r * per;
            v[r] = x[COLS ? n * lanes + lane : lane * N + n];
            if (r > 0 && Ns > 1) v[r] = cmulf(v[r], __ldg(tw + r * k * tstep));
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
            const int n = base_out + r * Ns;
            y[COLS ? n * lanes + lane : lane * N + n] = v[r];
        }
    }
}

// runs all stages; returns the buffer (0 or 1) that holds the result
template <bool COLS>
__device__ __forceinline__ int fft_in_smem(float2 *buf0, float2 *buf1, int N, int lanes, const float2 *tw) {
    int cur = 0;
    int Ns = 1;
    while (Ns * 4 <= N) {
        __syncthreads();
        if (cur == 0) stockham_stage<4, COLS>(buf0, buf1, N, Ns, lanes, tw);
        else stockham_stage<4, COLS>(buf1, buf0, N, Ns, lanes, tw);
        cur ^= 1;
        Ns *= 4;
    }
    if (Ns < N) {
        __syncthreads();
        if (cur == 0) stockham_stage<2, COLS>(buf0, buf1, N, Ns, lanes, tw);
        else stockham_stage<2, COLS>(buf1, buf0, N, Ns, lanes, tw);
        cur ^= 1;
    }
    __syncthreads();
    return cur;
}

struct FftArgs {
    const float2 *in[4];
    float2 *out[4];
    const float2 *tw;
    int ld_in, ld_out, N, other, lanes, in_roll_r, in_roll_c, out_roll;
};

// rows: `lanes` consecutive rows per CTA, transform along the contiguous axis
__global__ void __launch_bounds__(256) fft_rows_kernel(const FftArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    float2 *buf0 = fsm, *buf1 = fsm + (size_t)a.lanes * a.N;
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    const int row0 = blockIdx.x * a.lanes;
    const int total = a.lanes * a.N;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int lane = idx / a.N, n = idx - lane * a.N;
        const int r = row0 + lane;
        float2 v = make_float2(0.f, 0.f);
        if (r < a.other) {
            int rs = r - a.in_roll_r; if (rs < 0) rs += a.other;
            int cs = n - a.in_roll_c; if (cs < 0) cs += a.N;
            v = in[(size_t)rs * a.ld_in + cs];
        }
        buf0[idx] = v;
    }
    const int cur = fft_in_smem<false>(buf0, buf1, a.N, a.lanes, a.tw);
    const float2 *res = cur ? buf1 : buf0;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int lane = idx / a.N, n = idx - lane * a.N;       // n = position in the OUTPUT row
        const int r = row0 + lane;
        if (r < a.other) {
            int q = n - a.out_roll; if (q < 0) q += a.N;        // out[(q + roll) % N] = X[q]
            out[(size_t)r * a.ld_out + n] = res[lane * a.N + q];
        }
    }
}

// columns: `lanes` adjacent columns per CTA, transform along the strided axis
__global__ void __launch_bounds__(256) fft_cols_kernel(const FftArgs a) {
    extern __shared__ __align__(16) float2 fsm[];
    float2 *buf0 = fsm, *buf1 = fsm + (size_t)a.lanes * a.N;
    const float2 *__restrict__ in = pick4(a.in, blockIdx.y);
    float2 *__restrict__ out = pick4(a.out, blockIdx.y);
    const int c0 = blockIdx.x * a.lanes;
    const int total = a.lanes * a.N;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int n = idx / a.lanes, lane = idx - n * a.lanes;
        const int c = c0 + lane;
        buf0[idx] = (c < a.other) ? in[(size_t)n * a.ld_in + c] : make_float2(0.f, 0.f);
    }
    const int cur = fft_in_smem<true>(buf0, buf1, a.N, a.lanes, a.tw);
    const float2 *res = cur ? buf1 : buf0;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int n = idx / a.lanes, lane = idx - n * a.lanes;  // n = OUTPUT row
        const int c = c0 + lane;
        if (c < a.other) {
            int q = n - a.out_roll; if (q < 0) q += a.N;
            out[(size_t)n * a.ld_out + c] = res[q * a.lanes + lane];
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
constexpr size_t FFT_SMEM_BUDGET = 192 * 1024;

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

extern "C" int mlb_fft_max_length(void) { return (int)(mlb::FFT_SMEM_BUDGET / (2 * sizeof(float2))); }

extern "C" int mlb_fft_rows(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int n_rows,
                            int N, const mlb_c64 *tw, int in_roll_r, int in_roll_c, int out_roll, int batch,
                            void *stream) {
    mlb::FftArgs a;
    if (int rc = mlb::fill_args(a, h_in, h_out, batch, "mlb_fft_rows")) return rc;
    MLB_REQUIRE(mlb::is_pow2(N) && N >= 2 && N <= mlb_fft_max_length(), "mlb_fft_rows: length %d must be a power of two <= %d",
                N, mlb_fft_max_length());
    MLB_REQUIRE(tw && n_rows > 0 && ld_in >= N && ld_out >= N, "mlb_fft_rows: bad sizes");
    MLB_REQUIRE(in_roll_r >= 0 && in_roll_r < n_rows && in_roll_c >= 0 && in_roll_c < N && out_roll >= 0 && out_roll < N,
                "mlb_fft_rows: rolls out of range");
    for (int b = 0; b < batch; ++b)
        MLB_REQUIRE(in_roll_r == 0 || a.in[b] != a.out[b], "mlb_fft_rows: in-place needs in_roll_r == 0");
    a.tw = reinterpret_cast<const float2 *>(tw);
    a.ld_in = ld_in; a.ld_out = ld_out; a.N = N; a.other = n_rows;
    a.in_roll_r = in_roll_r; a.in_roll_c = in_roll_c; a.out_roll = out_roll;
    int lanes = 1024 / N; if (lanes < 1) lanes = 1; if (lanes > n_rows) lanes = n_rows;
    a.lanes = lanes;
    const size_t smem = 2 * (size_t)lanes * N * sizeof(float2);
    MLB_CUDA(cudaFuncSetAttribute(mlb::fft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlb::FFT_SMEM_BUDGET));
    dim3 grid((n_rows + lanes - 1) / lanes, batch);
    mlb::fft_rows_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
    return mlb::check_launch("mlb_fft_rows");
}

extern "C" int mlb_fft_cols(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int N,
                            int n_cols, const mlb_c64 *tw, int out_roll, int batch, void *stream) {
    mlb::FftArgs a;
    if (int rc = mlb::fill_args(a, h_in, h_out, batch, "mlb_fft_cols")) return rc;
    MLB_REQUIRE(mlb::is_pow2(N) && N >= 2 && N <= mlb_fft_max_length(), "mlb_fft_cols: length %d must be a power of two <= %d",
                N, mlb_fft_max_length());
    MLB_REQUIRE(tw && n_cols > 0 && ld_in >= n_cols && ld_out >= n_cols, "mlb_fft_cols: bad sizes");
    MLB_REQUIRE(out_roll >= 0 && out_roll < N, "mlb_fft_cols: roll out of range");
    a.tw = reinterpret_cast<const float2 *>(tw);
    a.ld_in = ld_in; a.ld_out = ld_out; a.N = N; a.other = n_cols;
    a.in_roll_r = 0; a.in_roll_c = 0; a.out_roll = out_roll;
    int lanes = (int)(mlb::FFT_SMEM_BUDGET / 2 / ((size_t)2 * N * sizeof(float2)));   // half the budget: 2 CTAs/SM
    if (lanes > 16) lanes = 16;
    if (lanes < 1) lanes = 1;
    if (lanes > n_cols) lanes = n_cols;
    a.lanes = lanes;
    const size_t smem = 2 * (size_t)lanes * N * sizeof(float2);
    MLB_CUDA(cudaFuncSetAttribute(mlb::fft_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mlb::FFT_SMEM_BUDGET));
    dim3 grid((n_cols + lanes - 1) / lanes, batch);
    mlb::fft_cols_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
    return mlb::check_launch("mlb_fft_cols");
}
