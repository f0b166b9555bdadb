// Error reporting, device queries and version of the C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace tq {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t err, const char* what) {
    if (err == cudaSuccess) return TQ_OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
    return err == cudaErrorMemoryAllocation ? TQ_ERR_OOM : TQ_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace tq

extern "C" int tq_version(void) { return 100; }
extern "C" const char* tq_last_error(void) { return tq::g_error; }
