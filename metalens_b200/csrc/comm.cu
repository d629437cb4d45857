// NCCL plumbing of the C-ABI (mlb_comm_init / mlb_allgather_P / mlb_allgather_fields /
// mlb_allreduce_scalar).  The reference has no collective call site; these are the exchange steps of 8e for
// callers that do not map peer memory (peer.cu is the NVLink peer-store path): one all-gather of the finished
// far-field power tiles (4*K^2/G bytes per rank), an all-gather of aperture slabs built by different ranks
// (32*M^2/G bytes per rank; nearfield.py:488-514 builds disjoint y slabs independently), and the scalar
// all-reduce of total_P / the incident power (nearfield_farfield.py:74, nearfield.py:474-477).
//
// libnccl is resolved at run time (dlopen of the copy already in the process -- torch ships one -- else the
// system one), so the library itself has no link-time dependency on NCCL and still loads on a box without it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"

namespace mlb {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return MLB_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("mlb_comm: libnccl.so.2 not found (%s)", dlerror());
        return MLB_ERR_UNSUPPORTED;
    }
#define MLB_NCCL_SYM(field, name)                                                        \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));             \
    if (!g_nccl.field) {                                                                 \
        set_error("mlb_comm: symbol %s missing in libnccl", name);                       \
        return MLB_ERR_UNSUPPORTED;                                                      \
    }
    MLB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    MLB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    MLB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    MLB_NCCL_SYM(AllGather, "ncclAllGather")
    MLB_NCCL_SYM(AllReduce, "ncclAllReduce")
    MLB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MLB_NCCL_SYM
    g_nccl.handle = h;
    return MLB_OK;
}

#define MLB_NCCL(call, what)                                                             \
    do {                                                                                 \
        ncclResult_t r_ = (call);                                                        \
        if (r_ != ncclSuccess) {                                                         \
            mlb::set_error("%s: NCCL error %s", what, mlb::g_nccl.GetErrorString(r_));   \
            return MLB_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

}  // namespace mlb

extern "C" int mlb_comm_unique_id(char *h_id128) {
    MLB_REQUIRE(h_id128, "mlb_comm_unique_id: NULL buffer");
    if (int rc = mlb::nccl_load()) return rc;
    static_assert(sizeof(ncclUniqueId) == MLB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    MLB_NCCL(mlb::g_nccl.GetUniqueId(&id), "mlb_comm_unique_id");
    memcpy(h_id128, &id, sizeof(id));
    return MLB_OK;
}

extern "C" int mlb_comm_init(int rank, int world, const char *h_id128, void **comm) {
    MLB_REQUIRE(h_id128 && comm && world >= 1 && rank >= 0 && rank < world, "mlb_comm_init: bad arguments");
    if (int rc = mlb::nccl_load()) return rc;
    ncclUniqueId id;
    memcpy(&id, h_id128, sizeof(id));
    ncclComm_t c = nullptr;
    MLB_NCCL(mlb::g_nccl.CommInitRank(&c, world, id, rank), "mlb_comm_init");
    *comm = c;
    return MLB_OK;
}

extern "C" int mlb_comm_destroy(void *comm) {
    if (!comm) return MLB_OK;
    if (int rc = mlb::nccl_load()) return rc;
    MLB_NCCL(mlb::g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(comm)), "mlb_comm_destroy");
    return MLB_OK;
}

extern "C" int mlb_allgather_P(void *comm, const float *send, float *recv, size_t count_per_rank, void *stream) {
    MLB_REQUIRE(comm && send && recv, "mlb_allgather_P: NULL argument");
    if (int rc = mlb::nccl_load()) return rc;
    MLB_NCCL(mlb::g_nccl.AllGather(send, recv, count_per_rank, ncclFloat, reinterpret_cast<ncclComm_t>(comm),
                                   (cudaStream_t)stream), "mlb_allgather_P");
    return MLB_OK;
}

extern "C" int mlb_allgather_fields(void *comm, const mlb_c64 *send, mlb_c64 *recv, size_t count_per_rank, void *stream) {
    MLB_REQUIRE(comm && send && recv, "mlb_allgather_fields: NULL argument");
    if (int rc = mlb::nccl_load()) return rc;
    MLB_NCCL(mlb::g_nccl.AllGather(send, recv, 2 * count_per_rank, ncclFloat, reinterpret_cast<ncclComm_t>(comm),
                                   (cudaStream_t)stream), "mlb_allgather_fields");
    return MLB_OK;
}

extern "C" int mlb_allreduce_scalar(void *comm, const double *send, double *recv, int n, void *stream) {
    MLB_REQUIRE(comm && send && recv && n >= 1, "mlb_allreduce_scalar: bad arguments");
    if (int rc = mlb::nccl_load()) return rc;
    MLB_NCCL(mlb::g_nccl.AllReduce(send, recv, (size_t)n, ncclDouble, ncclSum, reinterpret_cast<ncclComm_t>(comm),
                                   (cudaStream_t)stream), "mlb_allreduce_scalar");
    return MLB_OK;
}
