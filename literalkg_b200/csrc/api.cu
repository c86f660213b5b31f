// Library-level entry points: version, error string, device check.
#include "common.cuh"

namespace lkg {
namespace {
thread_local char g_error[512] = "";
}
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace lkg

extern "C" int lkg_abi_version(void) { return LKG_ABI_VERSION; }

extern "C" const char* lkg_last_error(void) { return lkg::g_error; }

extern "C" int lkg_device_check(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        cudaGetLastError();
        LKG_FAIL(LKG_ERR_CUDA, "cudaGetDeviceProperties(%d) failed: %s", device, cudaGetErrorString(e));
    }
    if (prop.major != 10)
        LKG_FAIL(LKG_ERR_ARCH, "device %d is sm_%d%d; liblkg is built for sm_100a only (no fallback)", device,
                 prop.major, prop.minor);
    return LKG_OK;
}
