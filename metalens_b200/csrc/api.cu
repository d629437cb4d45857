// Plumbing of the C-ABI: version, thread-local error string, device caps, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mlb {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace mlb

extern "C" {

int mlb_version(void) { return MLB_VERSION; }

const char *mlb_last_error(void) { return mlb::g_err; }

long long mlb_launch_count(void) { return mlb::g_launches.load(); }

int mlb_struct_sizes(int *out2) {
    MLB_REQUIRE(out2 != nullptr, "mlb_struct_sizes: out2 is NULL");
    out2[0] = (int)sizeof(mlb_table_pack);
    out2[1] = (int)sizeof(mlb_lens_desc);
    return MLB_OK;
}

int mlb_device_caps(int device, int *out4) {
    MLB_REQUIRE(out4 != nullptr, "mlb_device_caps: out4 is NULL");
    cudaDeviceProp p;
    MLB_CUDA(cudaGetDeviceProperties(&p, device));
    out4[0] = p.major;
    out4[1] = p.minor;
    out4[2] = p.multiProcessorCount;
    out4[3] = (int)p.sharedMemPerBlockOptin;
    return MLB_OK;
}

}  // extern "C"
