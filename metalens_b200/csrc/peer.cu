// Multi-GPU exchange steps of the far-field path over peer-mapped memory (NVLink 5 / NVSwitch), SURVEY 8e.
//
// The reference has no multi-device code; what it has is independence: disjoint uy chunks of one transform
// (nearfield_farfield.py:45-66) and disjoint y slabs of one assembly (nearfield.py:488-514).  The kernels here are
// the exchange steps that independence leaves over when the work is spread over GPUs:
//   * mlb_peer_allgather : every rank PUSHES its finished block (far-field power tiles, or the column slab of ONE
//     far field) into the same place of every peer's result buffer with plain 16-byte stores on peer-mapped
//     pointers.  Stores leave the SMs from L2-resident data, so nothing is read back over the links and the copy
//     engines stay free; a handful of CTAs next to the persistent HBM-bound row pass is enough to fill one NVLink
//     direction.  Two flag rows in symmetric memory replace the collective's barriers: ENTER (a rank may be
//     written to only after it has entered the same gather, i.e. after it is done with the buffer's previous
//     contents) and DONE (release-store after the last CTA's stores; mlb_peer_wait acquires it).
//   * mlb_peer_barrier   : flag barrier between the scattering row pass (fft.cu, mlb_fft_rows_scatter) and the
//     column pass of the distributed 2-D transform.
// Flags are monotonically increasing epochs kept in device memory (local_state), so that a captured CUDA graph
// can be replayed.  Every spin loop gives up after MLB_PEER_TIMEOUT_NS and raises local_state[3]: a lost peer
// must not hang the GPU.
#include "common.cuh"

namespace mlb {

constexpr unsigned long long PEER_TIMEOUT_NS = 20ULL * 1000ULL * 1000ULL * 1000ULL;
enum { FLAG_ENTER = 0, FLAG_DONE = 1, FLAG_BARRIER = 2 };
enum { ST_GATHER_EPOCH = 0, ST_GATHER_CTAS = 1, ST_BARRIER_EPOCH = 2, ST_ERROR = 3 };

struct PeerPtrs { void *p[MLB_MAX_PEERS]; };

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= epoch (wrap-safe signed distance); false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned int *flag, unsigned int epoch) {
    const unsigned long long t0 = now_ns();
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if ((++spins & 1023u) == 0 && now_ns() - t0 > PEER_TIMEOUT_NS) return false;
        __nanosleep(64);
    }
    return true;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerPtrs flags, int rank, int world,
                                                          unsigned int *__restrict__ state) {
    const unsigned int e = state[ST_BARRIER_EPOCH] + 1u;
    const int p = threadIdx.x;
    __threadfence_system();
    if (p < world) {
        st_release_sys(reinterpret_cast<unsigned int *>(flags.p[p]) + FLAG_BARRIER * MLB_MAX_PEERS + rank, e);
        if (!wait_flag(reinterpret_cast<const unsigned int *>(flags.p[rank]) + FLAG_BARRIER * MLB_MAX_PEERS + p, e))
            atomicExch(state + ST_ERROR, 1u);
    }
    __syncwarp();
    if (p == 0) state[ST_BARRIER_EPOCH] = e;
}

struct GatherArgs {
    const unsigned char *src;
    PeerPtrs dst, flags;
    long long src_pitch, dst_pitch, dst_offset;      // bytes
    int rows, row_vec;                               // row_vec = row_bytes / 16
    int rank, world;
    // small side block that travels with the gather (the rank's total_P block sums): n_aux doubles
    const double *aux_src;
    PeerPtrs aux_dst;
    int aux_offset, n_aux;
};

__global__ void __launch_bounds__(512) peer_allgather_kernel(const __grid_constant__ GatherArgs a,
                                                             unsigned int *__restrict__ state) {
    const unsigned int e = state[ST_GATHER_EPOCH] + 1u;
    const int tid = threadIdx.x;
    // ENTER: tell every peer this rank's result buffer may be overwritten, wait until every peer said the same
    if (tid < a.world) {
        st_release_sys(reinterpret_cast<unsigned int *>(a.flags.p[tid]) + FLAG_ENTER * MLB_MAX_PEERS + a.rank, e);
        if (!wait_flag(reinterpret_cast<const unsigned int *>(a.flags.p[a.rank]) + FLAG_ENTER * MLB_MAX_PEERS + tid, e))
            atomicExch(state + ST_ERROR, 1u);
    }
    __syncthreads();
    // push: every 16-byte vector is loaded once and stored to all peers (rotated order, so that at any moment the ranks
    // write to different destinations); U vectors per thread in flight -- the loop is latency-bound otherwise
    constexpr int U = 4;
    const long long per_peer = (long long)a.rows * a.row_vec;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + tid; base < per_peer; base += stride * U) {
        uint4 x[U];
        long long soff[U], doff[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long idx = base + u * stride;
            if (idx < per_peer) {
                const int row = (int)(idx / a.row_vec), v = (int)(idx - (long long)row * a.row_vec);
                soff[u] = (long long)row * a.src_pitch + (long long)v * 16;
                doff[u] = a.dst_offset + (long long)row * a.dst_pitch + (long long)v * 16;
                x[u] = *reinterpret_cast<const uint4 *>(a.src + soff[u]);
            } else {
                soff[u] = -1;
            }
        }
        for (int step = 0; step < a.world; ++step) {
            int p = a.rank + 1 + step;
            if (p >= a.world) p -= a.world;
            unsigned char *dbase = reinterpret_cast<unsigned char *>(a.dst.p[p]);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (soff[u] >= 0 && dbase + doff[u] != a.src + soff[u]) *reinterpret_cast<uint4 *>(dbase + doff[u]) = x[u];
        }
    }
    if (blockIdx.x == 0 && a.n_aux > 0) {
        for (int i = tid; i < a.n_aux * a.world; i += blockDim.x) {
            const int p = i / a.n_aux, k = i - p * a.n_aux;
            double *d = reinterpret_cast<double *>(a.aux_dst.p[p]) + a.aux_offset + k;
            if (d != a.aux_src + k) *d = a.aux_src[k];
        }
    }
    // DONE: the last CTA to finish publishes the epoch to every peer
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const unsigned int done = atomicAdd(state + ST_GATHER_CTAS, 1u);
        if (done == gridDim.x - 1) {
            state[ST_GATHER_CTAS] = 0u;
            state[ST_GATHER_EPOCH] = e;
            __threadfence_system();
            for (int p = 0; p < a.world; ++p)
                st_release_sys(reinterpret_cast<unsigned int *>(a.flags.p[p]) + FLAG_DONE * MLB_MAX_PEERS + a.rank, e);
        }
    }
}

__global__ void __launch_bounds__(32) peer_wait_kernel(const unsigned int *__restrict__ my_flags, int world,
                                                       unsigned int *__restrict__ state) {
    const unsigned int e = state[ST_GATHER_EPOCH];
    if ((int)threadIdx.x < world &&
        !wait_flag(my_flags + FLAG_DONE * MLB_MAX_PEERS + threadIdx.x, e))
        atomicExch(state + ST_ERROR, 1u);
}

// mlb_peer_wait followed by mlb_sum_f64 in one launch: the pushes have landed (block sums included), so total_P can be
// finished right here -- one kernel less on the critical path of the one-aperture step
__global__ void __launch_bounds__(256) peer_wait_sum_kernel(const unsigned int *__restrict__ my_flags, int world,
                                                            unsigned int *__restrict__ state, const double *__restrict__ in,
                                                            int n, double scale, double *__restrict__ out) {
    __shared__ double ws[8];
    const unsigned int e = state[ST_GATHER_EPOCH];
    if ((int)threadIdx.x < world && !wait_flag(my_flags + FLAG_DONE * MLB_MAX_PEERS + threadIdx.x, e))
        atomicExch(state + ST_ERROR, 1u);
    __syncthreads();
    const double s = ordered_sum_256(in, n, ws);
    if (threadIdx.x == 0) out[0] = s * scale;
}

static int fill_ptrs(PeerPtrs &out, void *const *h, int world, const char *who) {
    MLB_REQUIRE(h != nullptr && world >= 1 && world <= MLB_MAX_PEERS, "%s: world %d (1..%d)", who, world, MLB_MAX_PEERS);
    for (int p = 0; p < MLB_MAX_PEERS; ++p) out.p[p] = nullptr;
    for (int p = 0; p < world; ++p) {
        MLB_REQUIRE(h[p] != nullptr, "%s: peer pointer %d is NULL", who, p);
        out.p[p] = h[p];
    }
    return MLB_OK;
}

// CUDA loads kernels lazily, and loading one may have to wait for running kernels: a rank whose barrier kernel is
// already spinning must not be the one whose host thread then blocks loading the next exchange kernel (with virtual
// ranks on one device that is a deadlock, with real ranks a stall).  Force the loads up front.
static void preload_kernels() {
    static bool done = false;
    if (done) return;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, peer_barrier_kernel);
    cudaFuncGetAttributes(&fa, peer_allgather_kernel);
    cudaFuncGetAttributes(&fa, peer_wait_kernel);
    cudaFuncGetAttributes(&fa, peer_wait_sum_kernel);
    done = true;
}

}  // namespace mlb

extern "C" int mlb_peer_state_words(void) { return 8; }
extern "C" int mlb_peer_flag_words(void) {
    mlb::preload_kernels();
    return 3 * MLB_MAX_PEERS;
}

extern "C" int mlb_peer_barrier(void *const *h_flags, int rank, int world, void *local_state, void *stream) {
    mlb::PeerPtrs f;
    if (int rc = mlb::fill_ptrs(f, h_flags, world, "mlb_peer_barrier")) return rc;
    MLB_REQUIRE(rank >= 0 && rank < world && local_state, "mlb_peer_barrier: bad rank / state");
    mlb::peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, reinterpret_cast<unsigned int *>(local_state));
    return mlb::check_launch("mlb_peer_barrier");
}

extern "C" int mlb_peer_allgather(const void *src, long long src_pitch, int rows, long long row_bytes,
                                  void *const *h_dst, long long dst_pitch, long long dst_offset,
                                  const double *aux_src, void *const *h_aux_dst, int aux_offset, int n_aux,
                                  void *const *h_flags, int rank, int world, void *local_state, int n_ctas,
                                  void *stream) {
    mlb::GatherArgs a;
    if (int rc = mlb::fill_ptrs(a.dst, h_dst, world, "mlb_peer_allgather")) return rc;
    if (int rc = mlb::fill_ptrs(a.flags, h_flags, world, "mlb_peer_allgather")) return rc;
    MLB_REQUIRE(src && rank >= 0 && rank < world && local_state, "mlb_peer_allgather: bad rank / pointers");
    MLB_REQUIRE(rows > 0 && row_bytes > 0 && row_bytes % 16 == 0 && src_pitch % 16 == 0 && dst_pitch % 16 == 0 &&
                    dst_offset % 16 == 0 && src_pitch >= row_bytes && dst_pitch >= row_bytes,
                "mlb_peer_allgather: rows / pitches / offset must be positive multiples of 16 bytes");
    MLB_REQUIRE(mlb::aligned16(src), "mlb_peer_allgather: source not 16-byte aligned");
    for (int p = 0; p < world; ++p) MLB_REQUIRE(mlb::aligned16(h_dst[p]), "mlb_peer_allgather: destination %d not 16-byte aligned", p);
    MLB_REQUIRE(row_bytes / 16 < (1LL << 31), "mlb_peer_allgather: row too long");
    a.src = reinterpret_cast<const unsigned char *>(src);
    a.src_pitch = src_pitch; a.dst_pitch = dst_pitch; a.dst_offset = dst_offset;
    a.rows = rows; a.row_vec = (int)(row_bytes / 16); a.rank = rank; a.world = world;
    a.aux_src = aux_src; a.n_aux = 0; a.aux_offset = aux_offset;
    for (int p = 0; p < MLB_MAX_PEERS; ++p) a.aux_dst.p[p] = nullptr;
    if (aux_src && n_aux > 0) {
        if (int rc = mlb::fill_ptrs(a.aux_dst, h_aux_dst, world, "mlb_peer_allgather(aux)")) return rc;
        a.n_aux = n_aux;
    }
    if (n_ctas <= 0) n_ctas = 16;
    const long long work = ((long long)rows * a.row_vec + 511) / 512;
    if (n_ctas > work) n_ctas = (int)work;
    if (n_ctas > 148) n_ctas = 148;                 // all CTAs must be co-resident: they wait for each other's peers
    mlb::peer_allgather_kernel<<<n_ctas, 512, 0, (cudaStream_t)stream>>>(a, reinterpret_cast<unsigned int *>(local_state));
    return mlb::check_launch("mlb_peer_allgather");
}

extern "C" int mlb_peer_wait(const void *my_flags, int world, void *local_state, void *stream) {
    MLB_REQUIRE(my_flags && local_state && world >= 1 && world <= MLB_MAX_PEERS, "mlb_peer_wait: bad arguments");
    mlb::peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned int *>(my_flags), world,
                                                             reinterpret_cast<unsigned int *>(local_state));
    return mlb::check_launch("mlb_peer_wait");
}

extern "C" int mlb_peer_wait_sum(const void *my_flags, int world, void *local_state, const double *in, int n, double scale,
                                 double *out, void *stream) {
    MLB_REQUIRE(my_flags && local_state && world >= 1 && world <= MLB_MAX_PEERS && in && out && n >= 0,
                "mlb_peer_wait_sum: bad arguments");
    mlb::peer_wait_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned int *>(my_flags), world,
                                                                  reinterpret_cast<unsigned int *>(local_state), in, n, scale, out);
    return mlb::check_launch("mlb_peer_wait_sum");
}
