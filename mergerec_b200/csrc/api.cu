// api.cu -- version / error plumbing of the C ABI (include/mergerec_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mr {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}
}  // namespace mr

extern "C" int mr_version(void) { return 100; /* 0.1.0 */ }
extern "C" const char* mr_last_error(void) { return mr::g_err; }
