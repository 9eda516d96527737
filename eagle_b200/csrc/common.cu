// Error reporting and device queries of the C ABI.
#include "common.cuh"

#include <stdarg.h>

namespace egl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -(int)e;
}

}  // namespace egl

extern "C" int egl_version(void) { return EGL_ABI_VERSION; }

extern "C" const char* egl_last_error(void) { return egl::g_err; }

extern "C" int egl_build_flags(void) {
#ifdef EGL_BENCH_VARIANTS
    return 1;
#else
    return 0;
#endif
}

extern "C" int egl_sm_count(void) {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        egl::set_error("egl_sm_count: no CUDA device");
        return -1;
    }
    if (dev != cached_dev) {
        if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        cached_dev = dev;
    }
    return cached;
}
