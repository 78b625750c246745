// core.cu — library-level entry points: version, thread-local error string, device query.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

namespace g3d {
static thread_local char tls_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
    va_end(ap);
}

int sm_count(int device) {
    static std::atomic<int> cache[64];
    if (device < 0 || device >= 64) return 148;
    int n = cache[device].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        cache[device].store(n, std::memory_order_relaxed);
    }
    return n;
}
}  // namespace g3d

extern "C" const char* g3d_version(void) { return "geom3d-b200 0.1 (sm_100a)"; }

extern "C" const char* g3d_last_error(void) { return g3d::tls_error; }

extern "C" int g3d_sm_count(int device) {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        g3d::set_error("g3d_sm_count: %s", cudaGetErrorString(e));
        return G3D_ERR_CUDA;
    }
    return n;
}
