/*
 * metalens_b200 -- C-ABI of the B200-native near-field -> far-field engine.
 *
 * The reference (sbyrnes321/metalens) is pure Python; it has no FFI.  Its drop-in
 * boundary is the Python call surface (nearfield_farfield.farfield_from_nearfield,
 * nearfield.build_nearfield, GratingCollection/HexGridSet.build_interpolators).
 * This header is the compute boundary underneath that surface: the Python host
 * package `metalens_b200` binds exactly these symbols through ctypes
 * (metalens_b200/_lib.py) and nothing else.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name starts with `h_`;
 *    complex data is interleaved (re,im) float pairs ("c64"), 16-byte aligned,
 *    with an even leading dimension (in complex elements);
 *  - every function is asynchronous on `stream` (a cudaStream_t passed as void*);
 *  - return value 0 = OK, negative = error; mlb_last_error() gives the message
 *    (thread-local);
 *  - no hidden allocation: workspaces are passed in by the caller.
 */
#ifndef METALENS_B200_H
#define METALENS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLB_VERSION 100          /* 0.1.0 */
#define MLB_OK 0
#define MLB_ERR_ARG (-1)
#define MLB_ERR_CUDA (-2)
#define MLB_ERR_UNSUPPORTED (-3)

typedef struct mlb_c64 { float re, im; } mlb_c64;

/* ---- plumbing ------------------------------------------------------------ */
int mlb_version(void);
const char *mlb_last_error(void);
/* out[0]=sm major, [1]=sm minor, [2]=SM count, [3]=max dynamic smem per block (bytes) */
int mlb_device_caps(int device, int *out4);
/* number of kernels this library has launched since load (bench "gpu_launches") */
long long mlb_launch_count(void);
/* sizeof(mlb_table_pack), sizeof(mlb_lens_desc) as compiled into the library: a binding checks its own struct
 * layout against these before passing descriptors */
int mlb_struct_sizes(int *out2);

/* ---- A1/A4: separable aperture sum --------------------------------------- */
/*
 * Twiddle table  out[m*ld + i] = exp(i*pi*scale*coord[m]*u[i]),  m<n_coord, i<n_u,
 * phases formed in float64 and reduced exactly (sincospi) before rounding to fp32
 * (SURVEY H1).  With scale = -2*n_glass/wavelength, coord = x' and u = ux this is
 * the kernel e^{-ik x' ux} of nearfield_farfield.py:97-120; with scale = -2/N,
 * coord = 0..N-1 and u = integer bin numbers it is the exact DFT twiddle of
 * numpy.fft (nearfield_farfield.py:111-116).
 */
int mlb_twiddle_build(const double *coord, int n_coord, const double *u, int n_u,
                      double scale, mlb_c64 *out, int ld, void *stream);

/*
 * Tiled complex reduction (fp32 SIMT, 1-D TMA bulk copies into shared memory):
 *     C_b[r*ldc + c] = sum_{k<depth} At_b[k*lda + r] * B[k*ldb + c]      b < batch<=4
 * Both operands are stored with the contracted index as the ROW index, which is
 * how the aperture (J[m1][m2]), the twiddles and the stage-1 output naturally lie.
 *   stage 1:  At = field J_f (Mx x My),  B = AxT (Mx x Kx)  -> UT_f (My x Kx)
 *   stage 2:  At = UT_f (My x Kx),       B = Ay  (My x Ky)  -> Fhat_f (Kx x Ky)
 * Replaces the caller-side fft2(fftshift(.)) of nearfield_farfield.py:18-20 for an
 * arbitrary direction-cosine grid.
 */
int mlb_cgemm_tn(const mlb_c64 *const *h_At, int lda, const mlb_c64 *B, int ldb,
                 mlb_c64 *const *h_C, int ldc, int rows, int cols, int depth,
                 int batch, void *stream);

/*
 * Aperture fold for FFT-bin-stride grids: when the far-field grid is every s-th
 * FFT bin, the M-point sum collapses exactly to an (M/s)-point sum of the folded
 * aperture
 *     G_f[p1*ldg + p2] = sum_{t1<s1,t2<s2} J_f[((p1-h1) mod K1 + t1*K1)*ldj + (p2-h2) mod K2 + t2*K2]
 * with K1 = M1/s1, K2 = M2/s2 and h = the fftshift origin (M - M/2).  Streaming,
 * HBM-bound: reads 8*M1*M2 bytes per field once.
 */
int mlb_fold(const mlb_c64 *const *h_J, int ldj, int M1, int M2, int s1, int s2,
             int h1, int h2, mlb_c64 *const *h_G, int ldg, int batch, void *stream);

/* ---- A4 on the tensor cores: 3xTF32 complex GEMM (tcgen05 + TMEM + TMA), BASELINE cfg3 ---------- */
/*
 * Complex C = A.B embedded in a real K-major GEMM (see csrc/cgemm_tc.cu):
 *   A operand : complex64 row-major [rows][depth_c] viewed as fp32 [rows][2*depth_c]
 *   B operand : "embedding" fp32 [2*cols_c][2*depth_c]: row 2n = (Re,-Im) pairs of column n, row 2n+1 = (Im,Re)
 * Every operand comes as a tf32 hi/lo pair; D += Ah.Bh + Al.Bh + Ah.Bl in fp32 TMEM accumulators.
 * mlb_tf32_split   : fp32 matrix -> hi = tf32(x), lo = tf32(x - hi)   (aperture fields -> A operand)
 * mlb_twiddle_tf32 : exp(i*pi*scale*coord[m]*u[i]) from float64 phases, written as A operand
 *                    (layout 0: [n_u][2*n_coord]) or B embedding (layout 1: [2*n_u][2*n_coord])
 * mlb_cgemm_tc     : mode 1: result written as the B embedding (hi and lo) of the NEXT stage,
 *                            out[2j+q][2*row..2*row+1], pitch ldo floats;
 *                    mode 2: result written as complex64 out[row*ldo + j].
 *   NF->FF stage 1:  T[m1][j] = sum_m2 J[m1][m2] Ay[m2][j]  (A = split fields,  B = twiddle embedding, mode 1)
 *   NF->FF stage 2:  F[i][j]  = sum_m1 Ax[i][m1] T[m1][j]   (A = twiddle rows,  B = stage-1 output,    mode 2)
 * Pitches (floats) must be multiples of 4 (TMA needs 16-byte row pitches).
 */
int mlb_tf32_split(const float *in, int ld_in, float *hi, float *lo, int ld_out, int rows, int cols, void *stream);
int mlb_twiddle_tf32(const double *coord, int n_coord, const double *u, int n_u, double scale, int layout,
                     float *hi, float *lo, int ld, void *stream);
/* batch <= 4 problems of identical shape in one launch (host arrays of device pointers; repeat a
 * pointer to share an operand between items, e.g. the twiddles across the four fields) */
int mlb_cgemm_tc(const float *const *h_Ah, const float *const *h_Al, int lda, const float *const *h_Bh,
                 const float *const *h_Bl, int ldb, int rows, int cols_c, int depth_c, int mode,
                 float *const *h_out_hi, float *const *h_out_lo, int ldo, int batch, void *stream);

/* mlb_cgemm_tc with the contraction split into chunks of 512 real depth: every chunk starts its TMEM accumulators at zero
 * and its epilogue ADDS the partial sum to the complex64 result in memory with round-to-nearest.  The tensor core
 * accumulates with truncation -- a bias of ~2^-24 per 8-deep step that does not average out in a coherent sum (1e-4 at
 * depth 4096); chunking bounds it below 1e-5 whatever the aperture size (measured up to 2048^2 coherent apertures).  mode 1 needs h_scratch: one complex64
 * [rows][ld_scratch >= cols_c] buffer per batch item (the partial sums; embedded as the next stage's operand at the end);
 * mode 2 accumulates in h_out_hi itself (h_scratch may be NULL). */
int mlb_cgemm_tc_split(const float *const *h_Ah, const float *const *h_Al, int lda, const float *const *h_Bh,
                       const float *const *h_Bl, int ldb, int rows, int cols_c, int depth_c, int mode,
                       float *const *h_out_hi, float *const *h_out_lo, int ldo, int batch, mlb_c64 *const *h_scratch,
                       int ld_scratch, void *stream);

/* ---- A1 (FFT formulation): shared-memory FFT passes.  Lengths 2^a 3^b 5^c <= 8192 (the
 * sizes good_fft_number() produces); powers of two take the tuned kernels, others mixed-radix ones. ---- */
/* Twiddle tables for the FFT passes, 2*N entries: out[t] = exp(-2 pi i t / N) for t < N, followed by the
 * same values re-ordered per Stockham stage (conflict-free shared-memory reads); float64 phases rounded
 * once to fp32 */
int mlb_fft_twiddle(int N, mlb_c64 *out, void *stream);
/* Tuning knobs of the row pass (defaults are the B200-tuned values): loader variant (2 = TMA-fed
 * persistent producer/consumer kernel, the default for 256..2048 points; 1 = thread-issued loads, consecutive
 * samples per thread; 0 = first radix-4 stage done by the loader),
 * points per CTA, threads per CTA (64/128/256), vector width of the loads (1 or 2 complex). */
int mlb_fft_tune(int rows_plain_loader, int rows_points_per_cta, int rows_threads, int rows_vec);
/* Host only (no GPU): radix sequence the mixed engine uses for a length N = 2^a 3^b 5^c (good_fft_number() sizes,
 * nearfield.py:30-36): returns the stage count (< 0 on error), radix8[0..7] = radices (0-padded), *pad_shift = the
 * shared-memory padding rule (4 = one pad per 16 elements, 30 = none). */
int mlb_fft_mixed_plan(int N, int *radix8, int *pad_shift);
/* Host only: 1 if the register kernels of the mixed engine have an instantiation compiled for the plan of length N
 * (cols = 0: rows, one per CTA; cols = 1: a column (sub-)transform, 16 columns per CTA) and option mixed_compiled is on,
 * else 0 (the generic run-time-plan kernel serves it). */
int mlb_fft_mixed_compiled(int N, int cols);
/* longest transform the shared-memory passes support (8192 complex64) */
int mlb_fft_max_length(void);
/*
 * Batched 1-D DFT along rows (the contiguous axis) with the aperture fold and the fftshift
 * bookkeeping of nearfield_farfield.py:18-20, :68 folded into the indices.  Input matrices are
 * [n_rows*s1][N*s2]; the loader sums the s1*s2 aliased samples (see mlb_fold; s1 = s2 = 1 = no fold):
 *   G_b[r][p]  = sum_{t1<s1,t2<s2} in_b[((r - in_roll_r) mod n_rows) + t1*n_rows][((p - in_roll_c) mod N) + t2*N]
 *   out_b[r][(q + out_roll) % N] = sum_p G_b[r][p] e^{-2 pi i q p / N}
 * For a strided far-field grid this is the only kernel that touches the full aperture: it reads
 * 8*n_rows*s1*N*s2 bytes per field exactly once (HBM-bound).
 * transpose_out != 0 stores out_b[(q + out_roll) % N][r] instead (pitch ld_out >= n_rows): two such passes
 * make a 2-D transform with contiguous (TMA-streamed) reads in both; only where
 * mlb_fft_rows_can_transpose(N) returns 1 (TMA-fed kernel: 256..2048 points).
 */
int mlb_fft_rows_can_transpose(int N);
int mlb_fft_rows(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int n_rows, int N,
                 int s1, int s2, const mlb_c64 *tw, int in_roll_r, int in_roll_c, int out_roll, int transpose_out,
                 int batch, void *stream);
/* Same along columns:  out_b[(q + out_roll) % N][c] = sum_p in_b[p][c] e^{-2 pi i q p / N}.  In-place allowed
 * for power-of-two N <= 2048 and every other length; for power-of-two N >= 4096 the transform is a two-pass
 * decomposition that uses the INPUT buffer as scratch (it is overwritten) and needs out != in; for other lengths
 * the two-pass form is chosen when out != in and >= 8 columns (INPUT overwritten), the direct one otherwise. */
int mlb_fft_cols(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int N, int n_cols,
                 const mlb_c64 *tw, int out_roll, int batch, void *stream);

/*
 * Fused column pass + radiated-power epilogue (float32 P): the column DFT of all four row-pass outputs
 * in_f [N][ld_in] (f = Ex,Ey,Hx,Hy; same index convention as mlb_fft_cols) and, straight from registers,
 * P[(q + out_roll) % N][c] of mlb_ff_epilogue (nearfield_farfield.py:135-189; ux has N entries in the
 * OUTPUT row order, uy n_cols entries).  The 4 x N x n_cols aperture sums never go to memory unless h_Fhat
 * (4 device pointers, pitch ldf) is given.  block_sums receives mlb_fft_cols_power_blocks(N, n_cols) partial
 * sums of the finite P values (query it after any mlb_set_option call).  Powers of two 256..8192 with the
 * radix-16 column engine (cols_engine = 1; lengths >= 4096 run a first pass in place on in_f, which is therefore
 * scratch; h_Fhat must be NULL), 256..2048 with the radix-4 one: mlb_fft_cols_power_blocks returns 0 for every other
 * length and the caller uses mlb_fft_cols + mlb_ff_epilogue.
 */
int mlb_fft_cols_power_blocks(int N, int n_cols);
int mlb_fft_cols_power(const mlb_c64 *const *h_in, int ld_in, int N, int n_cols, const mlb_c64 *tw, int out_roll,
                       const double *ux, const double *uy, double amp_scale, double wavelength, double n_glass,
                       double Z0, float *P, int ldp, int accumulate, double *block_sums,
                       mlb_c64 *const *h_Fhat, int ldf, void *stream);
/* mlb_fft_cols_power + total_P in one call: total[0] = total_scale * (sum of the block sums), nearfield_farfield.py:74.
 * Where the pass is a single launch of 256-thread CTAs (radix-16 engine, 1024..2048-point columns) its last CTA adds the
 * block sums itself -- same fixed order as mlb_sum_f64, one launch less on the critical path of a pipelined step --,
 * otherwise mlb_sum_f64 follows.  done_counter: one zeroed uint32 of device memory per concurrently running call (left
 * zero again). */
int mlb_fft_cols_power_total(const mlb_c64 *const *h_in, int ld_in, int N, int n_cols, const mlb_c64 *tw, int out_roll,
                             const double *ux, const double *uy, double amp_scale, double wavelength, double n_glass,
                             double Z0, float *P, int ldp, int accumulate, double *block_sums, double *total,
                             double total_scale, void *done_counter, void *stream);
/* Named integer tuning options (defaults are the B200-tuned values):
 *   rows_ctas_per_sm     resident CTAs per SM of the TMA-fed row pass, 0 = as many as fit (default)
 *   rows_l2_evict_first  1 (default) = stream the aperture through L2 with an evict-first policy
 *   cols_power_wide      fused pass tile: 1 = 4096-point column tiles / 1024 threads, 0 = 2048 / 512,
 *                        -1 (default) = by length: wide from 1024 points up
 *   rows_ring_kb         shared-memory ring of the TMA-fed row pass per CTA: 64 (default) or 128
 *   rows_engine          0 = radix-4 shared-memory row kernels everywhere, 1 = radix-16 register kernels (256..8192
 *                        points) everywhere, 2 (default) = radix-16 without a fold, TMA-fed fold+FFT kernel with one
 *   cols_engine          0 = radix-4 column kernels, 1 (default) = radix-16 register kernels (256..8192 points; 4096 and
 *                        8192 as 16 x 256/512 in two passes, the first in place on the INPUT buffer)
 *   mixed_registers      non-power-of-two lengths (the good_fft_number() sizes; big-radix engine: radices up to 16 in registers,
 *                        long columns as A x B in two passes, the first in place on the INPUT buffer unless the call is
 *                        in place): 1 (default) = one-butterfly-per-thread register kernels where they pay (long
 *                        transforms that keep >= 80 % of a CTA's threads busy), 2 = wherever they apply, 0 = never
 *   mixed_compiled       1 (default) = those register kernels run the instantiation compiled for the plan (radices,
 *                        sub-lengths, lanes and padding as constants: no spills, half the registers; every plan the
 *                        dispatch sends there has one), 0 = the generic kernels that read the plan at run time
 *   mixed_occupancy      resident CTAs per SM those kernels are compiled for: 0 (default: compiled plans 4; generic rows
 *                        2, columns 4), 2..4 (compiled plans: 3 or 4)
 *   cols_strip_mb        two-pass (>= 4096-point) column transforms run strip by strip, strips of this many MB of all
 *                        fields (intermediate stays in L2); 0 (default) = one strip
 *   r16_min_lg           the radix-16 kernels serve lengths from 2^this up (default 10; 8..13), the radix-4 ones below
 *   r16_occupancy        resident CTAs per SM the radix-16 kernels are compiled for: 0 (default: rows 4, columns 3), 2..4
 * mlb_get_option returns -1 for an unknown name. */
int mlb_set_option(const char *name, int value);
int mlb_get_option(const char *name);

/* ---- A4 on a uniform "zoomed" grid: chirp-z (Bluestein) aperture sum on the FFT passes -----------------------
 * For ux_i = u0 + i du (any u0, du) and x_m = (m - origin) d the sum of nearfield_farfield.py:97-120 along one axis is
 *     F[i] = post[i] . conj( FFT_L( conj( FFT_L(pad(J . pre)) . FFT_L(kern) ) ) )[i + origin]        (L >= M + K - 1)
 * mlb_czt_chirps builds pre[M], kern[L] (the wrapped chirp e^{+i pi tq n^2/2}; its FFT is taken once with mlb_fft_rows) and
 * post[K] (includes 1/L) from t_lin = 2 n d u0 / lambda, t_quad = 2 n d du / lambda with float64 phases.
 * mlb_czt_pointwise is the pointwise step between the FFT passes (pad, multiply, conjugate, crop), batch <= 4 fields:
 *   out_b[r][c] = conj_out?( conj_in?(in_b[r + row_off][c + col_off]) . row_tab[r] . col_tab[c] ),  r < rows_valid and
 *   c < cols_valid, 0 elsewhere (r < rows_out, c < cols_out); row_tab / col_tab may be NULL. */
int mlb_czt_chirps(int M, int origin, int K, int L, double t_lin, double t_quad, mlb_c64 *pre, mlb_c64 *kern,
                   mlb_c64 *post, void *stream);
int mlb_czt_pointwise(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out, int ld_out, int rows_out, int cols_out,
                      int rows_valid, int cols_valid, int row_off, int col_off, const mlb_c64 *row_tab,
                      const mlb_c64 *col_tab, int conj_in, int conj_out, int batch, void *stream);

/* ---- A2/A3: radiated power ------------------------------------------------ */
/*
 * Fhat (4 x Kx x Ky c64: Ex,Ey,Hx,Hy aperture sums) -> P (Kx x Ky) following
 * farfield_from_nearfield_helper, nearfield_farfield.py:135-189, in float64
 * arithmetic: equivalent currents N,L; uz = sqrt(1-ux^2-uy^2) with NaN for
 * evanescent bins; theta/phi components with the 1e-9 regulariser; Cartesian
 * components at the exact ux==uy==0 bin; P = k^2/(32 pi^2 Z) (...)/(uz+1e-5) * 2.
 * `amp_scale` multiplies every Fhat before use (dx*dy, and 1/len factors).
 * P is written as float (p_is_double=0) or double (1); adding 2 to p_is_double ACCUMULATES into P instead
 * (incoherent sum over sources / polarisations, nearfield.py:69-73); adding 4 (float64 output only) says that h_Fhat
 * holds complex128 aperture sums (pitch ldf in complex128 elements) -- the strict drop-in's inputs as they are.  If `block_sums` is not
 * NULL, each block writes the sum of its finite P values to block_sums[blockIdx]
 * (mlb_ff_epilogue_blocks() entries) for a deterministic total_P.
 */
int mlb_ff_epilogue_blocks(int Kx, int Ky);
int mlb_ff_epilogue(const mlb_c64 *const *h_Fhat, int ldf, const double *ux, const double *uy,
                    int Kx, int Ky, double amp_scale, double wavelength, double n_glass,
                    double Z0, void *P, int ldp, int p_is_double, double *block_sums,
                    void *stream);
/*
 * A5 (figure of merit of a far field, new composition -- the reference's FOM lives in S4/Lua,
 * grating.lua:188-253): per-block sums (mlb_ff_epilogue_blocks() entries each) of P over the finite
 * bins inside the cone (ux-ux0)^2+(uy-uy0)^2 <= radius^2 and over all finite bins; reduce both
 * with mlb_sum_f64.  FOM = cone / total.
 */
int mlb_cone_power(const void *P, int ldp, int p_is_double, const double *ux, const double *uy, int Kx, int Ky,
                   double ux0, double uy0, double radius, double *cone_block_sums, double *total_block_sums,
                   void *stream);
/* out[0] = scale * sum_{i<n} in[i], summed in a fixed order by one block */
int mlb_sum_f64(const double *in, int n, double scale, double *out, void *stream);

/* ---- T1-T4 / B2-B7: amplitude tables and aperture-field assembly ---------------- */
#define MLB_MAX_PACKS 12

/* Flattened tables of one GratingCollection / HexGridSet (metalens_b200/tables.py):
 *   axes   : float64  ux[n_ux] | uy[n_uy] | third[n_g]     (third = grating period or index)
 *   values : complex128 [order][iu][iv][ig][slot], slot = 2*pol + amp
 *            (pol 0='x',1='y'; amp 0='ampfy',1='ampfx'), i.e. grating.py:1186-1232 /
 *            lens_center.py:188-226 tables with the four values one corner needs adjacent
 *   values_f32 : the same table as complex64 (float pairs), read by the complex64-output path
 *   orders : int32 [n_orders][2] = (ox,oy) in the reference's loop order (nearfield.py:264)
 *   order_map : int32 [(2R+1)^2], entry (ox+R)*(2R+1)+(oy+R) = index of (ox,oy) in `orders` or -1;
 *            R = order_radius = max |ox|,|oy|.  Lets the kernel visit only the orders that can propagate
 *   bounds : interpolator_bounds 6-tuple (grating.py:1230-1232)
 *   stats_slot : first slot of this pack in the `stats` array of mlb_nearfield_assemble */
typedef struct mlb_table_pack {
    const double *axes;
    const double *values;
    const float *values_f32;
    const int *orders;
    const int *order_map;
    int n_ux, n_uy, n_g, n_orders;
    double bounds[6];
    int stats_slot;
    int order_radius;
    /* uniform01 != 0: the ux and uy axes are uniformly spaced (what characterize() produces, grating.py:1167-1172):
     * the complex64-output kernel then locates the interpolation cell arithmetically, cell = (u - u_first) * inv_step,
     * instead of by bisection (same interpolant; an index may differ from searchsorted only where the weight is 0 or 1) */
    int uniform01, _pad;
    double u0_first, u0_inv_step, u1_first, u1_inv_step;
} mlb_table_pack;

/* Device-side packing of one collection's tables (SURVEY 8b "mlb_table_pack"; the name is taken by the struct): raw = the interpolators' value arrays as the
 * reference holds them (grating.py:1227-1229, lens_center.py:222-223), complex128 [order][slot][iu][iv][ig] with slot =
 * 2*pol + amp -> values (complex128) and values_f32 (complex64), both [order][iu][iv][ig][slot] as mlb_table_pack.values
 * / values_f32 expect. */
int mlb_table_pack_build(const double *raw, int n_orders, int n_ux, int n_uy, int n_g, double *values, float *values_f32,
                   void *stream);

/* Everything build_nearfield (nearfield.py:66-480) reads, as device arrays + scalars. */
typedef struct mlb_lens_desc {
    const double *x_pts, *y_pts;          /* sample coordinates, nx and ny entries          */
    int nx, ny;
    /* periphery rings, inside -> out (lens_periphery_summary, design_collimator.py:221-227) */
    const double *ring_boundary;          /* n_rings+1: r_min_list then lens_max_r (nearfield.py:125) */
    const double *r_center, *grating_period, *num_around;   /* n_rings each                 */
    const int *gc_index;                  /* n_rings: gratingcollection_index_here_list     */
    int n_rings, n_packs;
    mlb_table_pack packs[MLB_MAX_PACKS];  /* one per GratingCollection                      */
    /* centre cells (lens_center_summary rows x,y,index), bucketed into a uniform bin grid   */
    const double *cell_x, *cell_y;        /* n_cells, sorted by bin                         */
    const int *cell_which, *cell_orig;    /* grating index; original row number (exact distance ties
                                             go to the highest row, cKDTree's usual choice)   */
    const int *bin_start;                 /* nbx*nby+1 offsets into the sorted cells        */
    int n_cells, nbx, nby, _pad;
    double bin_x0, bin_y0, bin_size;
    mlb_table_pack hex;                   /* HexGridSet tables                              */
    double hex_x_period, hex_y_period;    /* nearfield.py:391-392                           */
    /* source (nearfield.py:66-73): point dipole at (sx,sy,sz<0) or plane wave (plane_wave=1) */
    double source_x, source_y, source_z;
    int plane_wave, source_pol;           /* pol: 0='x' 1='y' 2='z'                         */
    double wavelength, n_glass, dipole_moment, c0, Z0;
    /* Per-lens derived data: caller-allocated device buffers that mlb_nearfield_prepare() fills once per lens and
     * wavelength (they do not depend on the source or the sample grid); mlb_nearfield_assemble() reads them.
     *   ring_aux     n_rings x 64 bytes: { r_center, grating_period, angle_per_grating = 2 pi / num_around (:161),
     *                lateral_period (:165), 2 pi / grating_period, 2 pi / lateral_period, t2 (float64 each),
     *                gc index, i2 (int32 each) } -- (i2, t2) = interpolation cell and weight of the ring's
     *                grating_period on the third table axis (scipy RGI find_indices), one 64-byte record per ring
     *   ring_aux_f32 n_rings x 32 bytes: float32 screens for the order box and the grating-copy index
     *   ring_lut     n_lut + 1 ints: ring_lut[b] = number of ring boundaries < b * lut_r_max / n_lut (last entry
     *                n_rings + 1), so that searchsorted(boundaries, r) (:125) only looks at a 3-bin bracket;
     *                n_lut >= 1, about 4 * n_rings recommended
     *   lut_r_max    host copy of ring_boundary[n_rings] (lens_max_r, :94) */
    void *ring_aux;
    void *ring_aux_f32;
    int *ring_lut;
    int n_lut, _pad2;
    double lut_r_max;
    /*   ring_tables  n_rings x ring_table_stride floats (16-byte aligned; stride = mlb_nearfield_ring_table_floats(), a
     *                multiple of 4): per ring the complex64 slice [order][iu][iv][slot] of its collection's tables
     *                interpolated at the ring's grating period (scipy's linear rule on the third axis, float64, rounded
     *                once) -- the complex64-output kernel then needs 4 corners per order instead of 8 */
    float *ring_tables;
    long long ring_table_stride;
} mlb_lens_desc;

#define MLB_STATS_PER_ORDER 8   /* count, min/max ux, min/max uy, min/max third, pad (int64 each) */

/*
 * Fused build_nearfield: one thread per aperture sample computes ring lookup, incident dipole
 * field, grating-frame rotation, the diffraction-order loop with trilinear table gathers
 * (nearfield.py:263-327), propagation phase, rotation back, and the centre region with a
 * nearest-cell search (:359-466); all in float64.  Output fields are written as complex64
 * (out_is_double=0) or complex128 (1), row pitch `ld` elements.  power_block_sums receives
 * mlb_nearfield_blocks() partial sums of Ex_inc*Hy_inc - Ey_inc*Hx_inc over lens points
 * (:474-477; one partial sum per warp).  violation[0] is set non-zero if any interpolation point lies outside a pack's
 * bounds (the ValueErrors of :294-305, :412-419); with want_stats=1 the per-(pack,order)
 * count / min / max needed for the reference's messages are accumulated in `stats`
 * (int64, order-preserving encoding of doubles; see metalens_b200/nearfield.py).
 */
int mlb_nearfield_blocks(int nx, int ny);
/* Fills ring_aux / ring_aux_f32 / ring_lut of the descriptor (device buffers owned by the caller) from its ring
 * arrays and table axes; once per lens and wavelength, before the first mlb_nearfield_assemble(). */
int mlb_nearfield_prepare(const mlb_lens_desc *h_desc, void *stream);
/* floats per ring of mlb_lens_desc.ring_tables for this descriptor's packs (host only) */
long long mlb_nearfield_ring_table_floats(const mlb_lens_desc *h_desc);
/* tuning knob: register budget variant of the complex64 kernel (min resident blocks/SM: 1, 5, 6 or 8), or
 * 100 + lg: warp tile of 2^lg samples along y times 32 / 2^lg along x (lg 2..5; query mlb_nearfield_blocks after) */
int mlb_nearfield_tune(int min_blocks);
int mlb_nearfield_assemble(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld,
                           int out_is_double, double *power_block_sums, long long *stats, int want_stats,
                           int *violation, void *stream);
/* Samples whose INDEX choice hinges on the last bit of a library call in the reference:
 *  (1) exact nearest-cell ties (nearfield.py:363-364): a centre sample exactly equidistant from two hex cells gets the
 *      highest original row here, while the reference's cKDTree returns whichever cell its traversal meets first;
 *  (2) periphery samples within 1e-9 of the boundary between two grating copies: round(phi / angle_per_grating)
 *      (nearfield.py:167-169) flips with one ulp of arctan2.
 * mlb_nearfield_assemble_ties is mlb_nearfield_assemble that also reports those samples: tie_count[0] (zero it first)
 * counts them and tie_list receives the first `tie_capacity` linear sample indices i*ny + j.  A binding that wants the
 * reference's choice computes it on the host with the reference's own calls (cKDTree.query; numpy arctan2 / round) and
 * re-assembles just those samples with mlb_nearfield_fixup: sample fix_samples[t] takes fix_cells[t] = the SORTED
 * position of its cell (index into cell_x / cell_y) if it lies in the centre region, = the grating-copy index
 * round(phi / angle_per_grating) if it lies on a ring; the power sums are not touched. */
int mlb_nearfield_assemble_ties(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld,
                                int out_is_double, double *power_block_sums, long long *stats, int want_stats,
                                int *violation, int *tie_count, int *tie_list, int tie_capacity, void *stream);
int mlb_nearfield_fixup(const mlb_lens_desc *h_desc, void *Ex, void *Ey, void *Hx, void *Hy, int ld, int out_is_double,
                        const int *fix_samples, const int *fix_cells, int n_fix, int *violation, void *stream);

/* Trilinear gather of ONE table (scipy RegularGridInterpolator linear mode):
 * axes = u0[n0]|u1[n1]|u2[n2], values complex128 [n0][n1][n2], pts float64 [n][3] -> out complex128 [n] */
int mlb_table_eval(const double *axes, int n0, int n1, int n2, const double *values, const double *pts, int n,
                   double *out, void *stream);

/* ---- N2: lens layout on the device (design_collimator.py:74-137, lens_center.py:175-186) ----------------------
 * Hex-lattice centre of a design: lattice points n*(n2*sqrt(3)/2, n1 + n2/2) with x^2 + y^2 < radius^2 over the candidate
 * ranges of design_collimator.hexagonal_grid (:85-103), in the reference's row order (n2 outer, n1 inner).
 *   mlb_hex_count : counts[c] = cells of lattice column n2_lo + c                         (n2_hi - n2_lo + 1 ints)
 *   mlb_hex_fill  : offsets = exclusive scan of counts (int64); writes cells[row] = (x, y, index), index = the
 *                   HexGridSet entry picked by pick_from_phase(target_phase(r) + pi): argmax_k Im(x_amp[k] e^{-i phase});
 *                   x_amp = n_amp complex128 values (re, im pairs)
 * mlb_cells_bin : the bin grid the assembly kernel searches (mlb_lens_desc.cell_x / cell_y / cell_which / cell_orig /
 *   bin_start) from device cells [n][3]: phase 0 adds the bin populations into count_or_cursor (nbx*nby ints, zeroed by
 *   the caller), phase 1 scatters the cells with count_or_cursor = a copy of bin_start used as write cursors. */
int mlb_hex_count(double pitch, double radius, int n1_lo, int n1_hi, int n2_lo, int n2_hi, int *counts, void *stream);
int mlb_hex_fill(double pitch, double radius, int n1_lo, int n1_hi, int n2_lo, int n2_hi, const long long *offsets,
                 double wavelength, double refractive_index, double source_distance, const double *x_amp, int n_amp,
                 double *cells, void *stream);
int mlb_cells_bin(const double *cells, int n, double x0, double y0, double bin_size, int nbx, int nby, int phase,
                  int *count_or_cursor, double *cell_x, double *cell_y, int *cell_which, int *cell_orig, void *stream);

/* ---- 8e: multi-GPU exchange steps ------------------------------------------------------------------------------
 * The reference is single-process; its independent units are the uy chunks of one transform
 * (nearfield_farfield.py:45-66) and the y slabs of one assembly (nearfield.py:488-514).  Two families:
 *  (1) peer-memory kernels (NVLink 5 / NVSwitch, plain ld/st on peer-mapped pointers).  The caller maps the peers'
 *      buffers (CUDA IPC / torch symmetric memory) and passes host arrays of `world` device pointers, entry p = the
 *      buffer of rank p as mapped in THIS process (entry `rank` = the local one);
 *  (2) NCCL wrappers (mlb_comm_*) for callers without peer mappings.                                              */
#define MLB_MAX_PEERS 16
#define MLB_COMM_ID_BYTES 128
/* sizes (in 32-bit words) of the two bookkeeping blocks of the peer kernels: `flags` -- one block per rank in
 * peer-mapped (symmetric) memory, zero before first use -- and `local_state` -- plain device memory of the calling rank,
 * zero before first use; word 3 of local_state becomes non-zero if a wait ever timed out (20 s: a lost peer must not
 * hang the GPU).  One (flags, local_state) pair per stream of exchanges. */
int mlb_peer_flag_words(void);
int mlb_peer_state_words(void);
/* Row pass of ONE 2-D transform spread over `world` ranks, fused with the all-to-all that follows it: like mlb_fft_rows
 * (fold + fftshift rolls included) on this rank's n_rows folded rows -- input matrices [n_rows*s1][N*s2], folded row r
 * = sum over t1 of input rows r + t1*n_rows, i.e. the rank holds the s1 aliased copies of ITS rows one after the other.
 * The rank's rows are blocks of row_block consecutive rows of the distributed intermediate, row_stride apart (blocks dealt
 * round-robin to the ranks balance the work of a round lens; one block = a contiguous slab): output row r is row
 * (out_row0 + (r / row_block) * row_stride + r % row_block) mod n_rows_total, and its column slab
 * [p*N/world, (p+1)*N/world) is stored straight into rank p's buffer h_out_peers[p*batch + f] (pitch ld_out >= N/world
 * complex) over NVLink while the next rows are still streaming in.  After a mlb_peer_barrier every rank holds all
 * n_rows_total rows of its N/world columns and runs mlb_fft_cols / mlb_fft_cols_power on them.
 * world and row_block must be powers of two (world <= MLB_MAX_PEERS dividing N); N in 256..2048 (the TMA-fed kernel),
 * else MLB_ERR_UNSUPPORTED. */
int mlb_fft_rows_scatter(const mlb_c64 *const *h_in, int ld_in, mlb_c64 *const *h_out_peers, int ld_out, int n_rows,
                         int N, int s1, int s2, const mlb_c64 *tw, int in_roll_c, int out_roll, int out_row0,
                         int row_block, int row_stride, int n_rows_total, int world, int batch, void *stream);
/* Flag barrier across the ranks, in stream order: returns (on the device) once every rank's stream has reached its
 * matching mlb_peer_barrier -- all peer stores issued by earlier kernels of those streams are then visible. */
int mlb_peer_barrier(void *const *h_flags, int rank, int world, void *local_state, void *stream);
/* The ONE all-gather at the end, as pushes: copies the local block src (rows x row_bytes, pitch src_pitch)
 * to byte offset dst_offset of every rank's destination h_dst[p] (pitch dst_pitch; a destination that IS the source is
 * skipped), plus, optionally, n_aux doubles (the rank's total_P block sums) to h_aux_dst[p] + aux_offset.  Peers are
 * written only after they entered the same gather (their previous result may be overwritten); completion is published
 * with a release store that mlb_peer_wait acquires.  n_ctas (0 = default 16) CTAs of 512 threads; 16-byte granularity. */
int mlb_peer_allgather(const void *src, long long src_pitch, int rows, long long row_bytes, void *const *h_dst,
                       long long dst_pitch, long long dst_offset, const double *aux_src, void *const *h_aux_dst,
                       int aux_offset, int n_aux, void *const *h_flags, int rank, int world, void *local_state,
                       int n_ctas, void *stream);
/* Wait (on the device, in stream order) until every peer's pushes of the latest mlb_peer_allgather this rank ran have
 * landed in this rank's destination.  my_flags = this rank's flag block. */
int mlb_peer_wait(const void *my_flags, int world, void *local_state, void *stream);
/* mlb_peer_wait followed by mlb_sum_f64(in, n, scale, out) in one launch (total_P of the gathered block sums) */
int mlb_peer_wait_sum(const void *my_flags, int world, void *local_state, const double *in, int n, double scale,
                      double *out, void *stream);

/* NCCL wrappers (libnccl resolved at run time; MLB_ERR_UNSUPPORTED if it cannot be found).  mlb_comm_unique_id on one
 * rank, ship the MLB_COMM_ID_BYTES bytes to the others by any means (the launcher's store, MPI, a file), then
 * mlb_comm_init on every rank with its CUDA device current. */
int mlb_comm_unique_id(char *h_id128);
int mlb_comm_init(int rank, int world, const char *h_id128, void **comm);
int mlb_comm_destroy(void *comm);
/* recv[p*count + i] = send_p[i]: far-field power tiles (float32) / aperture slabs (complex64) */
int mlb_allgather_P(void *comm, const float *send, float *recv, size_t count_per_rank, void *stream);
int mlb_allgather_fields(void *comm, const mlb_c64 *send, mlb_c64 *recv, size_t count_per_rank, void *stream);
/* recv[i] = sum over ranks of send[i], float64 (total_P, incident power); in place allowed */
int mlb_allreduce_scalar(void *comm, const double *send, double *recv, int n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* METALENS_B200_H */
