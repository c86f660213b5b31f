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

// ---- peer push: one kernel stores a block of local rows into the copies of up to 8 peers over NVLink ----------------
// The row-partitioned path exchanges activation tables by PUSHING a rank's row block into every peer's copy (symmetric
// memory, literalkg_b200/parallel.py).  The copy engines move most of it; this kernel moves the rest on a few SMs at the
// same time so that both paths load the NVLinks together (one device-to-device copy per peer at a time reaches ~610 of
// the ~770 GB/s a GPU can send).  Every 16-byte vector is loaded once and stored to all destinations.
namespace lkg {
namespace {
struct PushDst {
    void* p[8];
};
__global__ void __launch_bounds__(256) peer_push_kernel(const uint4* __restrict__ src, int64_t n_vec, PushDst dst,
                                                        int n_dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k * stride < n_vec) v[k] = __ldg(src + i + k * stride);
        for (int d = 0; d < n_dst; ++d) {
            uint4* out = reinterpret_cast<uint4*>(dst.p[d]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i + k * stride < n_vec) out[i + k * stride] = v[k];
        }
    }
}
}  // namespace
}  // namespace lkg

extern "C" int lkg_peer_push(const void* src, int64_t nbytes, void* const* dst, int32_t n_dst, int32_t n_ctas,
                             void* stream_) {
    LKG_REQUIRE(src && dst && nbytes >= 0 && n_dst >= 1 && n_dst <= 8 && n_ctas >= 1, "bad push arguments");
    LKG_REQUIRE(nbytes % 16 == 0 && lkg::aligned16(src), "push blocks must be 16-byte aligned and sized");
    if (nbytes == 0) return LKG_OK;
    lkg::PushDst d{};
    for (int i = 0; i < n_dst; ++i) {
        LKG_REQUIRE(dst[i] && lkg::aligned16(dst[i]), "destination %d is null or not 16-byte aligned", i);
        d.p[i] = dst[i];
    }
    lkg::peer_push_kernel<<<(unsigned)n_ctas, 256, 0, (cudaStream_t)stream_>>>(static_cast<const uint4*>(src), nbytes / 16,
                                                                            d, n_dst);
    LKG_LAUNCH_CHECK("peer_push_kernel");
    return LKG_OK;
}
