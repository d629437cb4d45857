// Tiled complex reduction, fp32 SIMT:  C[r][c] = sum_k At[k][r] * B[k][c]   (complex64).
//
// This is the aperture sum of nearfield_farfield.py:97-120 in separable form:
// both stages of  Fhat = Ax . J . Ay  are instances of this kernel because the aperture, the
// twiddle tables and the stage-1 output all lie with the contracted index as the row index.
//
// Structure (B200): one CTA owns a BM x BN output tile; each k-slab of BK rows of At and B is a set
// of contiguous 1-KB row segments, so it is staged into shared memory with 1-D TMA bulk copies
// (cp.async.bulk -> UBLKCP) signalled through an mbarrier ring of STAGES slabs; 256 threads each
// keep a TM x TN complex micro-tile in registers (fp32 accumulate).  The micro-tile is strided in
// units of 32 so that every LDS.128 of a warp is either a broadcast (At) or 256 contiguous bytes
// (B): no bank conflicts, and global stores are 256-byte contiguous per half-warp.
#include "common.cuh"

namespace mlb {

struct CgemmArgs {
    const float2 *At[4];
    float2 *C[4];
    const float2 *B;
    int lda, ldb, ldc, rows, cols, depth;
};

template <int TM, int TN, int BK, int STAGES>
struct CgemmCfg {
    static constexpr int BM = 16 * TM, BN = 16 * TN;
    static constexpr int STAGE_ELEMS = BK * (BM + BN);
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * sizeof(float2) + STAGES * sizeof(uint64_t);
};

template <int TM, int TN>
__device__ __forceinline__ void cmac_tile(float2 (&acc)[TM][TN], const float4 (&a)[TM / 2], const float4 (&b)[TN / 2]) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const float ar = (i & 1) ? a[i >> 1].z : a[i >> 1].x;
        const float ai = (i & 1) ? a[i >> 1].w : a[i >> 1].y;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const float br = (j & 1) ? b[j >> 1].z : b[j >> 1].x;
            const float bi = (j & 1) ? b[j >> 1].w : b[j >> 1].y;
            acc[i][j].x = fmaf(ar, br, acc[i][j].x);
            acc[i][j].x = fmaf(-ai, bi, acc[i][j].x);
            acc[i][j].y = fmaf(ar, bi, acc[i][j].y);
            acc[i][j].y = fmaf(ai, br, acc[i][j].y);
        }
    }
}

template <int TM, int TN, int BK, int STAGES>
__global__ void __launch_bounds__(256, 1) cgemm_tn_kernel(const CgemmArgs args) {
    using Cfg = CgemmCfg<TM, TN, BK, STAGES>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 *tiles = reinterpret_cast<float2 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)STAGES * Cfg::STAGE_ELEMS * sizeof(float2));

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
    const float2 *__restrict__ At = pick4(args.At, blockIdx.z);
    const float2 *__restrict__ B = args.B;

    // how many elements of each smem row are really copied (never past the row pitch)
    const int na = min(BM, args.lda - r0), nb = min(BN, args.ldb - c0);
    const int nslab = (args.depth + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int slab) {  // executed by warp 0
        const int stage = slab % STAGES;
        const int k0 = slab * BK;
        const int kv = min(BK, args.depth - k0);
        float2 *as = tiles + (size_t)stage * Cfg::STAGE_ELEMS;
        float2 *bs = as + BK * BM;
        if (tid == 0) mbar_expect_tx(&full[stage], (uint32_t)(kv * (na + nb) * sizeof(float2)));
        __syncwarp();
        for (int i = tid; i < 2 * BK; i += 32) {
            const int kk = i % BK;
            if (kk < kv) {
                if (i < BK) bulk_g2s(as + kk * BM, At + (size_t)(k0 + kk) * args.lda + r0, na * 8u, &full[stage]);
                else bulk_g2s(bs + kk * BN, B + (size_t)(k0 + kk) * args.ldb + c0, nb * 8u, &full[stage]);
            }
        }
    };

    if (tid < 32) {
        for (int s = 0; s < STAGES - 1 && s < nslab; ++s) issue(s);
    }

    float2 acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = make_float2(0.f, 0.f);

    for (int slab = 0; slab < nslab; ++slab) {
        // the stage being refilled was consumed in iteration slab-1 (trailing __syncthreads)
        if (tid < 32 && slab + STAGES - 1 < nslab) issue(slab + STAGES - 1);
        const int stage = slab % STAGES;
        mbar_wait(&full[stage], (uint32_t)((slab / STAGES) & 1));
        const float2 *as = tiles + (size_t)stage * Cfg::STAGE_ELEMS + 2 * ty;
        const float2 *bs = tiles + (size_t)stage * Cfg::STAGE_ELEMS + BK * BM + 2 * tx;
        const int kv = min(BK, args.depth - slab * BK);
        if (kv == BK) {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float4 a[TM / 2], b[TN / 2];
#pragma unroll
                for (int j = 0; j < TM / 2; ++j) a[j] = *reinterpret_cast<const float4 *>(as + kk * BM + 32 * j);
#pragma unroll
                for (int j = 0; j < TN / 2; ++j) b[j] = *reinterpret_cast<const float4 *>(bs + kk * BN + 32 * j);
                cmac_tile<TM, TN>(acc, a, b);
            }
        } else {
            for (int kk = 0; kk < kv; ++kk) {
                float4 a[TM / 2], b[TN / 2];
#pragma unroll
                for (int j = 0; j < TM / 2; ++j) a[j] = *reinterpret_cast<const float4 *>(as + kk * BM + 32 * j);
#pragma unroll
                for (int j = 0; j < TN / 2; ++j) b[j] = *reinterpret_cast<const float4 *>(bs + kk * BN + 32 * j);
                cmac_tile<TM, TN>(acc, a, b);
            }
        }
        __syncthreads();
    }

    float2 *__restrict__ C = pick4(args.C, blockIdx.z);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = r0 + 32 * (i >> 1) + 2 * ty + (i & 1);
        if (r >= args.rows) continue;
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) {
            const int c = c0 + 32 * j + 2 * tx;
            float2 *dst = C + (size_t)r * args.ldc + c;
            if (c + 1 < args.cols) {
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x,
                                                               acc[i][2 * j + 1].y);
            } else if (c < args.cols) {
                *dst = acc[i][2 * j];
            }
        }
    }
}

template <int TM, int TN, int BK, int STAGES>
static int launch_cgemm(const CgemmArgs &a, int batch, cudaStream_t stream) {
    using Cfg = CgemmCfg<TM, TN, BK, STAGES>;
    auto kern = cgemm_tn_kernel<TM, TN, BK, STAGES>;
    static unsigned long long attr_set = 0; const unsigned long long devbit_ = mlb::device_bit();
    if (!(attr_set & devbit_)) {
        MLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set |= devbit_;
    }
    dim3 grid((a.cols + Cfg::BN - 1) / Cfg::BN, (a.rows + Cfg::BM - 1) / Cfg::BM, batch);
    kern<<<grid, 256, Cfg::SMEM, stream>>>(a);
    return check_launch("mlb_cgemm_tn");
}

}  // namespace mlb

extern "C" int mlb_cgemm_tn(const mlb_c64 *const *h_At, int lda, const mlb_c64 *B, int ldb, mlb_c64 *const *h_C,
                            int ldc, int rows, int cols, int depth, int batch, void *stream) {
    MLB_REQUIRE(h_At && B && h_C, "mlb_cgemm_tn: NULL pointer");
    MLB_REQUIRE(batch >= 1 && batch <= 4, "mlb_cgemm_tn: batch %d not in 1..4", batch);
    MLB_REQUIRE(rows > 0 && cols > 0 && depth > 0, "mlb_cgemm_tn: empty problem (%d,%d,%d)", rows, cols, depth);
    MLB_REQUIRE(lda >= rows && ldb >= cols && ldc >= cols, "mlb_cgemm_tn: leading dimension too small");
    MLB_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0,
                "mlb_cgemm_tn: leading dimensions must be even (16-byte rows): %d %d %d", lda, ldb, ldc);
    mlb::CgemmArgs a;
    a.B = reinterpret_cast<const float2 *>(B);
    MLB_REQUIRE(mlb::aligned16(B), "mlb_cgemm_tn: B not 16-byte aligned");
    for (int b = 0; b < 4; ++b) {
        const int s = b < batch ? b : 0;
        MLB_REQUIRE(h_At[s] && h_C[s], "mlb_cgemm_tn: NULL operand in batch slot %d", s);
        MLB_REQUIRE(mlb::aligned16(h_At[s]) && mlb::aligned16(h_C[s]), "mlb_cgemm_tn: operand %d not 16-byte aligned", s);
        a.At[b] = reinterpret_cast<const float2 *>(h_At[s]);
        a.C[b] = reinterpret_cast<float2 *>(h_C[s]);
    }
    a.lda = lda; a.ldb = ldb; a.ldc = ldc; a.rows = rows; a.cols = cols; a.depth = depth;
    // Large problems: 128x128 tiles (8x8 complex per thread).  Small ones: 64x64 tiles so the
    // grid still covers the 148 SMs.
    const long long big_tiles = (long long)((rows + 127) / 128) * ((cols + 127) / 128) * batch;
    if (big_tiles >= 148) return mlb::launch_cgemm<8, 8, 16, 4>(a, batch, (cudaStream_t)stream);
    return mlb::launch_cgemm<4, 4, 16, 4>(a, batch, (cudaStream_t)stream);
}
