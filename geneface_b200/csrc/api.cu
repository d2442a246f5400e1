// Error plumbing + misc entry points of libgfrender.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "gf_common.cuh"

namespace gf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return GF_ERR_CUDA;
    }
    return GF_OK;
}

}  // namespace gf

extern "C" {

GF_API const char* gf_last_error(void) { return gf::g_err; }

GF_API int gf_version(void) { return 100; }

GF_API int gf_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return p.major == 10 ? 1 : 0;
}

}  // extern "C"
