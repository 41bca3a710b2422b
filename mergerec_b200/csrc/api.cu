// api.cu -- version / error plumbing of the C ABI (include/mergerec_b200.h).
#include <math.h>
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

// Host-side helper of the metric finalisation (no CUDA involved): the value of CPython's builtin `sum()` over a list of
// floats -- the reference's `sum(ndcgs) / len(ndcgs)` (evaluator/metrics.py:88).  CPython >= 3.12 adds floats with
// Neumaier's compensated summation (Python/bltinmodule.c, builtin_sum_impl); older interpreters add them plainly.
// compensated != 0 selects the former.  Bit-identical to the interpreter's own result (asserted in the tests).
extern "C" double mr_float_sum(const double* x, int64_t n, int compensated) {
    if (!x || n <= 0) return 0.0;
    double f_result = x[0];
    if (!compensated) {
        for (int64_t i = 1; i < n; ++i) f_result += x[i];
        return f_result;
    }
    double c = 0.0;
    for (int64_t i = 1; i < n; ++i) {
        const double v = x[i];
        const double t = f_result + v;
        if (fabs(f_result) >= fabs(v)) c += (f_result - t) + v;
        else c += (v - t) + f_result;
        f_result = t;
    }
    if (c != 0.0 && isfinite(c)) f_result += c;
    return f_result;
}
