// Error channel / device helpers for libmvsnerf_b200_probes.so (the product library has its own in api.cu).
#include "../common.cuh"

namespace mvsn {
static thread_local char g_probe_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_probe_err, sizeof(g_probe_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    return n;
}
}  // namespace mvsn

extern "C" const char* mvsn_probe_last_error(void) { return mvsn::g_probe_err; }
