// Library-wide plumbing of the C ABI: version, thread-local error string, device info.
#include "common.cuh"

#include <string.h>

namespace mb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
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

}  // namespace mb

MB_API int mb_version(void) { return 100; }  // 0.1.0 -> major*10000 + minor*100 + patch

MB_API const char* mb_last_error(void) { return mb::g_err; }

MB_API int mb_sm_count(void) { return mb::sm_count(); }
