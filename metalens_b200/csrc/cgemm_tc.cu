// Tensor-core variant of the separable aperture sum (BASELINE cfg3 "dense exp(ik.r) tensor-core
// path"): complex GEMM on the 5th-generation tensor cores with fp32-class accuracy ("3xTF32").
//
//   complex  C = A . B   is embedded in a real GEMM (no wasted MACs):
//     A_real[m][2k+p]           = (Re, Im) interleaved  -> a complex64 row-major matrix AS IS
//     B_emb[2n+0][2k..2k+1]     = ( Re B[k][n], -Im B[k][n])
//     B_emb[2n+1][2k..2k+1]     = ( Im B[k][n],  Re B[k][n])
//     D[m][2n+q] = sum_kappa A_real[m][kappa] * B_emb[2n+q][kappa]  = (Re C[m][n], Im C[m][n])
//   Both operands are K-major, i.e. exactly the layout TMA + UMMA descriptors (SWIZZLE_128B) want.
//
//   TF32 keeps 10 mantissa bits; the 1e-5 parity target needs more, so every operand is split into
//   hi = tf32(x), lo = tf32(x - hi) and   D = sum Ah.Bh  +  sum (Al.Bh + Ah.Bl)   (the dropped Al.Bl
//   term is 2^-22 relative).  The tensor core adds into its fp32 accumulator with truncation, so the
//   large main term and the 2^-11-times-smaller correction term get SEPARATE TMEM accumulators
//   (2 x 256 columns = all of TMEM) and are summed in fp32 by the epilogue: the correction products no
//   longer cost a truncation of the big accumulator each (3x fewer biased roundings).
//   Truncation is a BIAS (toward zero, ~2^-24 of the accumulator per 8-deep step) that does not average out in a
//   coherent sum such as the focus of a lens: ~1e-4 at depth 4096.  The contraction is therefore split into chunks
//   of TC_CHUNK_BLOCKS k-blocks (64 accumulation steps): every launch starts its TMEM accumulators at zero and its
//   epilogue ADDS the chunk into the fp32 result in memory with round-to-nearest (unbiased) -- the in-chunk bias is
//   bounded whatever the aperture size, at the price of a launch, a pipeline fill and an epilogue (read-modify-write of
//   the output tile) per chunk.  Measured on coherent 1024^2 / 2048^2 lens apertures: 16-block chunks meet 1e-5,
//   32-block chunks do not (and are 1.45x faster at cfg3: 2.0e8 vs 1.4e8 points/s) -- accuracy decides.
//
// Kernel anatomy (one 128 x 256 real output tile per CTA, 192 threads):
//   warp 0 : TMA producer  (cp.async.bulk.tensor 2-D, 4 operand tiles per k-block, mbarrier tx)
//   warp 1 : TMEM allocation + single-thread tcgen05.mma issue (12 UMMAs per k-block), tcgen05.commit
//   warps 2-5 : epilogue, tcgen05.ld 32 lanes x 16 columns at a time -> global
//
// Stage 1 of NF->FF:  T[m1][j]  = sum_m2 J[m1][m2] . Ay[m2][j]   (A = aperture, B_emb = twiddles);
//                     the epilogue writes T directly as the B_emb operand of stage 2 (hi and lo).
// Stage 2          :  F[i][j]   = sum_m1 Ax[i][m1] . T[m1][j]    (A = twiddles,  B_emb = stage-1 output);
//                     the epilogue writes complex64 F[i][j].
#include <cuda.h>

#include "common.cuh"

namespace mlb {

constexpr int TC_BM = 128;       // real rows of D per CTA  (= UMMA M)
constexpr int TC_BN = 256;       // real cols of D per CTA  (= UMMA N) = 128 complex columns
constexpr int TC_BK = 32;        // tf32 elements per k-block = 128 bytes = one swizzle span
constexpr int TC_STAGES = 2;
constexpr int TC_CHUNK_BLOCKS = 16;   // k-blocks per launch (512 real depth = 64 truncating accumulation steps)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;        // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;        // 32 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // Ah, Al, Bh, Bl = 96 KB
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 192;

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(tf32_rna(x));
    lo = __uint_as_float(tf32_rna(x - hi));
}

// ---- tcgen05 / TMA wrappers -------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c_inner, int c_outer, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, rows of 128 bytes packed densely:
// start>>4 | LBO=1 (ignored for swizzled K-major) | SBO = 1024 B (8 rows) | version 1 | layout 2
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, N=256, M=128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

struct alignas(64) TcMaps {   // per batch item: A hi, A lo, B hi, B lo
    CUtensorMap m[4][4];
};

struct TcArgs {
    float *out_hi[4], *out_lo[4];   // mode 1: embedded operand hi / lo;  mode 2: out_hi = complex64 result
    int ldo;                  // floats (mode 1) or complex elements (mode 2)
    int rows, cols_c;         // valid output rows (real M) and complex columns (N/2)
    int kb0, k_blocks;        // this launch contracts k-blocks [kb0, kb0 + k_blocks)
    int mode;                 // 1: embedded hi/lo operand of the next stage, 2: complex64 result, 3: complex64 partial sum
    int accumulate;           // modes 2 / 3: add to what is in memory (round-to-nearest) instead of overwriting
};

__global__ void __launch_bounds__(TC_THREADS, 1)
cgemm_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs a) {
    const CUtensorMap *mapAh = &maps.m[blockIdx.z][0], *mapAl = &maps.m[blockIdx.z][1];
    const CUtensorMap *mapBh = &maps.m[blockIdx.z][2], *mapBl = &maps.m[blockIdx.z][3];
    extern __shared__ unsigned char tc_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(base + TC_STAGES * TC_STAGE_BYTES);
    uint64_t *empty = full + TC_STAGES;
    uint64_t *tmem_full = empty + TC_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM;          // first real row of the tile
    const int n0 = blockIdx.x * TC_BN;          // first real column (= row of the B operand)

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * TC_BN);      // main + correction accumulators: 512 columns x 128 lanes
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < a.k_blocks; ++kb) {
                const int s = kb % TC_STAGES;
                if (kb >= TC_STAGES) mbar_wait(&empty[s], (uint32_t)(((kb / TC_STAGES) - 1) & 1));
                unsigned char *st = base + s * TC_STAGE_BYTES;
                mbar_expect_tx(&full[s], TC_STAGE_BYTES);
                const int k0 = (a.kb0 + kb) * TC_BK;
                tma_load_2d(st, mapAh, k0, m0, &full[s]);
                tma_load_2d(st + TC_A_BYTES, mapAl, k0, m0, &full[s]);
                tma_load_2d(st + 2 * TC_A_BYTES, mapBh, k0, n0, &full[s]);
                tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, mapBl, k0, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            for (int kb = 0; kb < a.k_blocks; ++kb) {
                const int s = kb % TC_STAGES;
                mbar_wait(&full[s], (uint32_t)((kb / TC_STAGES) & 1));
                tc_fence_after();
                const uint32_t st = smem_u32(base + s * TC_STAGE_BYTES);
                const uint64_t dAh = make_desc_sw128(st), dAl = make_desc_sw128(st + TC_A_BYTES);
                const uint64_t dBh = make_desc_sw128(st + 2 * TC_A_BYTES), dBl = make_desc_sw128(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {          // UMMA K = 8 tf32 = 32 bytes -> +2 in the >>4 address field
                    const uint64_t adv = (uint64_t)(kk * 2);
                    umma_tf32(tmem_d, dAh + adv, dBh + adv, TC_IDESC, (kb | kk) != 0);           // main term
                    umma_tf32(tmem_d + TC_BN, dAl + adv, dBh + adv, TC_IDESC, (kb | kk) != 0);   // corrections
                    umma_tf32(tmem_d + TC_BN, dAh + adv, dBl + adv, TC_IDESC, 1);
                }
                umma_commit(&empty[s]);                           // smem slot free once these MMAs have read it
            }
            umma_commit(tmem_full);                               // accumulator complete
        }
    } else {
        // ------------------------------------------------ epilogue warps (2..5): TMEM -> registers -> global
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int lane_base = (warp & 3) * 32;                    // TMEM lanes this warp may touch
        float *const out_hi = pick4(a.out_hi, blockIdx.z), *const out_lo = pick4(a.out_lo, blockIdx.z);
        const int row = m0 + lane_base + lane;                    // real output row held by this thread
        for (int c = 0; c < TC_BN; c += 16) {
            uint32_t v[16], w[16];
            tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)c, v);
            tmem_ld16(tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)(TC_BN + c), w);
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = __float_as_uint(__uint_as_float(v[q]) + __uint_as_float(w[q]));
            if (row < a.rows) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int jc = (n0 + c) / 2 + q;              // complex column
                    if (jc >= a.cols_c) break;
                    const float re = __uint_as_float(v[2 * q]), im = __uint_as_float(v[2 * q + 1]);
                    if (a.mode == 1) {
                        // operand embedding for stage 2: rows 2j, 2j+1 of [2*cols_c][ldo], columns 2*row, 2*row+1
                        float rh, rl, ih, il;
                        split_tf32(re, rh, rl);
                        split_tf32(im, ih, il);
                        const size_t o0 = (size_t)(2 * jc) * a.ldo + 2 * row, o1 = o0 + a.ldo;
                        *reinterpret_cast<float2 *>(out_hi + o0) = make_float2(rh, -ih);
                        *reinterpret_cast<float2 *>(out_hi + o1) = make_float2(ih, rh);
                        *reinterpret_cast<float2 *>(out_lo + o0) = make_float2(rl, -il);
                        *reinterpret_cast<float2 *>(out_lo + o1) = make_float2(il, rl);
                    } else {
                        float2 *dst = reinterpret_cast<float2 *>(out_hi) + (size_t)row * a.ldo + jc;
                        if (a.accumulate) {                           // chunk sums are added with round-to-nearest
                            const float2 old = *dst;
                            *dst = make_float2(old.x + re, old.y + im);
                        } else {
                            *dst = make_float2(re, im);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 2 * TC_BN);
}

// ---- operand preparation kernels ----------------------------------------------------------------------
__global__ void tf32_split_kernel(const float *__restrict__ in, size_t ld_in, float *__restrict__ hi,
                                  float *__restrict__ lo, size_t ld_out, int rows, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= cols || r >= rows) return;
    float h, l;
    split_tf32(in[(size_t)r * ld_in + c], h, l);
    hi[(size_t)r * ld_out + c] = h;
    lo[(size_t)r * ld_out + c] = l;
}

// complex64 T[row][j] (pitch ld_in complex) -> the hi / lo operand embedding of the next stage (what the mode-1 epilogue
// writes directly when the contraction is a single chunk): rows 2j, 2j+1 of [2*cols_c][ldo], columns 2*row, 2*row+1
__global__ void tc_embed_kernel(const float2 *__restrict__ in, size_t ld_in, float *__restrict__ hi, float *__restrict__ lo,
                                size_t ldo, int rows, int cols_c) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;       // fast index = output column pair
    const int jc = blockIdx.y;
    if (row >= rows || jc >= cols_c) return;
    const float2 v = in[(size_t)row * ld_in + jc];
    float rh, rl, ih, il;
    split_tf32(v.x, rh, rl);
    split_tf32(v.y, ih, il);
    const size_t o0 = (size_t)(2 * jc) * ldo + 2 * row, o1 = o0 + ldo;
    *reinterpret_cast<float2 *>(hi + o0) = make_float2(rh, -ih);
    *reinterpret_cast<float2 *>(hi + o1) = make_float2(ih, rh);
    *reinterpret_cast<float2 *>(lo + o0) = make_float2(rl, -il);
    *reinterpret_cast<float2 *>(lo + o1) = make_float2(il, rl);
}

// twiddle w(u_i, x_m) = exp(i*pi*scale*coord[m]*u[i]) in float64, written hi/lo as
//   layout 0 (A operand):   row i, columns (2m, 2m+1) = (Re, Im)                      [n_u][ld]
//   layout 1 (B embedding): row 2i   columns (2m, 2m+1) = (Re, -Im),
//                           row 2i+1 columns (2m, 2m+1) = (Im,  Re)                   [2 n_u][ld]
__global__ void twiddle_tf32_kernel(const double *__restrict__ coord, int n_coord, const double *__restrict__ u, int n_u,
                                    double scale, int layout, float *__restrict__ hi, float *__restrict__ lo, size_t ld) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (m >= n_coord || i >= n_u) return;
    double s, c;
    sincospi(scale * coord[m] * u[i], &s, &c);
    const float ch = __uint_as_float(tf32_rna((float)c)), sh = __uint_as_float(tf32_rna((float)s));
    const float cl = __uint_as_float(tf32_rna((float)(c - (double)ch))), sl = __uint_as_float(tf32_rna((float)(s - (double)sh)));
    if (layout == 0) {
        const size_t o = (size_t)i * ld + 2 * m;
        *reinterpret_cast<float2 *>(hi + o) = make_float2(ch, sh);
        *reinterpret_cast<float2 *>(lo + o) = make_float2(cl, sl);
    } else {
        const size_t o0 = (size_t)(2 * i) * ld + 2 * m, o1 = o0 + ld;
        *reinterpret_cast<float2 *>(hi + o0) = make_float2(ch, -sh);
        *reinterpret_cast<float2 *>(hi + o1) = make_float2(sh, ch);
        *reinterpret_cast<float2 *>(lo + o0) = make_float2(cl, -sl);
        *reinterpret_cast<float2 *>(lo + o1) = make_float2(sl, cl);
    }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D fp32 tensor [rows][k] with row pitch ld floats; box = [box_rows][32 floats], 128-byte swizzle, OOB -> 0
static int make_map(CUtensorMap *map, const float *ptr, int rows, int k, int ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    MLB_REQUIRE(enc != nullptr, "mlb_cgemm_tc: cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MLB_REQUIRE(r == CUDA_SUCCESS, "mlb_cgemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%d k=%d ld=%d", (int)r, rows, k, ld);
    return MLB_OK;
}

}  // namespace mlb

extern "C" int mlb_tf32_split(const float *in, int ld_in, float *hi, float *lo, int ld_out, int rows, int cols,
                              void *stream) {
    MLB_REQUIRE(in && hi && lo && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= cols, "mlb_tf32_split: bad arguments");
    dim3 grid((cols + 255) / 256, rows);
    mlb::tf32_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ld_in, hi, lo, ld_out, rows, cols);
    return mlb::check_launch("mlb_tf32_split");
}

extern "C" int mlb_twiddle_tf32(const double *coord, int n_coord, const double *u, int n_u, double scale, int layout,
                                float *hi, float *lo, int ld, void *stream) {
    MLB_REQUIRE(coord && u && hi && lo && n_coord > 0 && n_u > 0, "mlb_twiddle_tf32: bad arguments");
    MLB_REQUIRE((layout == 0 || layout == 1) && ld >= 2 * n_coord && ld % 4 == 0,
                "mlb_twiddle_tf32: layout must be 0/1 and ld a multiple of 4 floats >= 2*n_coord");
    dim3 grid((n_coord + 127) / 128, n_u);
    mlb::twiddle_tf32_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(coord, n_coord, u, n_u, scale, layout, hi, lo, (size_t)ld);
    return mlb::check_launch("mlb_twiddle_tf32");
}

static int cgemm_tc_impl(const float *const *h_Ah, const float *const *h_Al, int lda, const float *const *h_Bh,
                         const float *const *h_Bl, int ldb, int rows, int cols_c, int depth_c, int mode,
                         float *const *h_out_hi, float *const *h_out_lo, int ldo, int batch, mlb_c64 *const *h_scratch,
                         int ld_scratch, void *stream) {
    MLB_REQUIRE(h_Ah && h_Al && h_Bh && h_Bl && h_out_hi, "mlb_cgemm_tc: NULL pointer");
    MLB_REQUIRE(batch >= 1 && batch <= 4, "mlb_cgemm_tc: batch %d not in 1..4", batch);
    MLB_REQUIRE(mode == 1 || mode == 2, "mlb_cgemm_tc: mode must be 1 (embedded hi/lo output) or 2 (complex64 output)");
    MLB_REQUIRE(mode == 2 || h_out_lo, "mlb_cgemm_tc: mode 1 needs out_lo");
    MLB_REQUIRE(rows > 0 && cols_c > 0 && depth_c > 0, "mlb_cgemm_tc: empty problem");
    const int K = 2 * depth_c;                       // real depth
    MLB_REQUIRE(lda >= K && ldb >= K && lda % 4 == 0 && ldb % 4 == 0, "mlb_cgemm_tc: operand pitch must be >= 2*depth and a multiple of 4 floats");
    MLB_REQUIRE(mode == 2 ? ldo >= cols_c : (ldo >= 2 * rows && ldo % 2 == 0), "mlb_cgemm_tc: output pitch too small");
    mlb::TcMaps maps;
    mlb::TcArgs a;
    for (int b = 0; b < 4; ++b) {
        const int s = b < batch ? b : 0;
        MLB_REQUIRE(h_Ah[s] && h_Al[s] && h_Bh[s] && h_Bl[s] && h_out_hi[s] && (mode == 2 || h_out_lo[s]),
                    "mlb_cgemm_tc: NULL operand in batch slot %d", s);
        MLB_REQUIRE(mlb::aligned16(h_Ah[s]) && mlb::aligned16(h_Al[s]) && mlb::aligned16(h_Bh[s]) && mlb::aligned16(h_Bl[s]),
                    "mlb_cgemm_tc: operands not 16-byte aligned");
        if (int rc = mlb::make_map(&maps.m[b][0], h_Ah[s], rows, K, lda, mlb::TC_BM)) return rc;
        if (int rc = mlb::make_map(&maps.m[b][1], h_Al[s], rows, K, lda, mlb::TC_BM)) return rc;
        if (int rc = mlb::make_map(&maps.m[b][2], h_Bh[s], 2 * cols_c, K, ldb, mlb::TC_BN)) return rc;
        if (int rc = mlb::make_map(&maps.m[b][3], h_Bl[s], 2 * cols_c, K, ldb, mlb::TC_BN)) return rc;
        a.out_hi[b] = h_out_hi[s];
        a.out_lo[b] = (mode == 1) ? h_out_lo[s] : nullptr;
    }
    static unsigned long long attr_set = 0; const unsigned long long devbit_ = mlb::device_bit();
    if (!(attr_set & devbit_)) {
        MLB_CUDA(cudaFuncSetAttribute(mlb::cgemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mlb::TC_SMEM));
        attr_set |= devbit_;
    }
    a.rows = rows; a.cols_c = cols_c;
    const int k_blocks = (K + mlb::TC_BK - 1) / mlb::TC_BK;
    dim3 grid((2 * cols_c + mlb::TC_BN - 1) / mlb::TC_BN, (rows + mlb::TC_BM - 1) / mlb::TC_BM, batch);
    const cudaStream_t st = (cudaStream_t)stream;
    const bool chunked = h_scratch != nullptr && k_blocks > mlb::TC_CHUNK_BLOCKS;
    if (!chunked) {                                       // short contraction (or no scratch): one launch, as before
        a.ldo = ldo; a.kb0 = 0; a.k_blocks = k_blocks; a.mode = mode; a.accumulate = 0;
        mlb::cgemm_tc_kernel<<<grid, mlb::TC_THREADS, mlb::TC_SMEM, st>>>(maps, a);
        return mlb::check_launch("mlb_cgemm_tc");
    }
    // chunked: every launch contracts TC_CHUNK_BLOCKS k-blocks from zero and adds its partial sum to the complex64
    // result in memory (round-to-nearest); mode 1 accumulates in the scratch and embeds it for the next stage at the end
    mlb::TcArgs c = a;
    if (mode == 1) {
        MLB_REQUIRE(ld_scratch >= cols_c, "mlb_cgemm_tc_split: scratch pitch too small");
        for (int b = 0; b < 4; ++b) {
            const int s = b < batch ? b : 0;
            MLB_REQUIRE(h_scratch[s] != nullptr, "mlb_cgemm_tc_split: NULL scratch %d", s);
            c.out_hi[b] = reinterpret_cast<float *>(const_cast<mlb_c64 *>(h_scratch[s]));
            c.out_lo[b] = nullptr;
        }
        c.ldo = ld_scratch; c.mode = 3;
    } else {
        c.ldo = ldo; c.mode = 2;
    }
    for (int kb0 = 0; kb0 < k_blocks; kb0 += mlb::TC_CHUNK_BLOCKS) {
        c.kb0 = kb0;
        c.k_blocks = (k_blocks - kb0 < mlb::TC_CHUNK_BLOCKS) ? k_blocks - kb0 : mlb::TC_CHUNK_BLOCKS;
        c.accumulate = kb0 > 0;
        mlb::cgemm_tc_kernel<<<grid, mlb::TC_THREADS, mlb::TC_SMEM, st>>>(maps, c);
        if (int rc = mlb::check_launch("mlb_cgemm_tc(chunk)")) return rc;
    }
    if (mode == 1) {
        for (int b = 0; b < batch; ++b) {
            dim3 ge((rows + 127) / 128, cols_c);
            mlb::tc_embed_kernel<<<ge, 128, 0, st>>>(reinterpret_cast<const float2 *>(h_scratch[b]), (size_t)ld_scratch,
                                                     h_out_hi[b], h_out_lo[b], (size_t)ldo, rows, cols_c);
            if (int rc = mlb::check_launch("mlb_cgemm_tc(embed)")) return rc;
        }
    }
    return MLB_OK;
}

extern "C" int mlb_cgemm_tc(const float *const *h_Ah, const float *const *h_Al, int lda, const float *const *h_Bh,
                            const float *const *h_Bl, int ldb, int rows, int cols_c, int depth_c, int mode,
                            float *const *h_out_hi, float *const *h_out_lo, int ldo, int batch, void *stream) {
    return cgemm_tc_impl(h_Ah, h_Al, lda, h_Bh, h_Bl, ldb, rows, cols_c, depth_c, mode, h_out_hi, h_out_lo, ldo, batch,
                         nullptr, 0, stream);
}

extern "C" int mlb_cgemm_tc_split(const float *const *h_Ah, const float *const *h_Al, int lda, const float *const *h_Bh,
                                  const float *const *h_Bl, int ldb, int rows, int cols_c, int depth_c, int mode,
                                  float *const *h_out_hi, float *const *h_out_lo, int ldo, int batch,
                                  mlb_c64 *const *h_scratch, int ld_scratch, void *stream) {
    MLB_REQUIRE(mode == 2 || h_scratch, "mlb_cgemm_tc_split: mode 1 needs a complex64 scratch per batch item");
    static mlb_c64 *const dummy[4] = {nullptr, nullptr, nullptr, nullptr};
    return cgemm_tc_impl(h_Ah, h_Al, lda, h_Bh, h_Bl, ldb, rows, cols_c, depth_c, mode, h_out_hi, h_out_lo, ldo, batch,
                         h_scratch ? h_scratch : dummy, ld_scratch, stream);
}
