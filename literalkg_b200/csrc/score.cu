// predict_links post-processing (model.py:488-491): global min-max normalisation + threshold -> int32.
#include "common.cuh"

namespace lkg {
namespace {

__global__ void minmax_reset_kernel(uint32_t* mm) {
    mm[0] = 0xffffffffu;
    mm[1] = 0u;
}

__global__ void threshold_kernel(const float* __restrict__ s, int64_t lds, int64_t rows, int64_t cols,
                                 const uint32_t* __restrict__ mm, float milestone, int32_t* __restrict__ pred,
                                 int64_t ldp) {
    auto dec = [](uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); };
    const float lo = dec(mm[0]), hi = dec(mm[1]);
    const float range = hi - lo;
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const float v = (s[r * lds + c] - lo) / range;   // same op order as model.py:490; 0/0 -> NaN -> 0
        pred[r * ldp + c] = v > milestone ? 1 : 0;
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_minmax_reset(uint32_t* minmax_dev, void* stream_) {
    LKG_REQUIRE(minmax_dev != nullptr, "minmax is null");
    minmax_reset_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(minmax_dev);
    LKG_LAUNCH_CHECK("minmax_reset_kernel");
    return LKG_OK;
}

extern "C" int lkg_predict_threshold(const float* scores, int64_t ld_scores, int64_t n_heads, int64_t n_tails,
                                     const uint32_t* minmax_dev, float milestone, int32_t* pred, int64_t ld_pred,
                                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(scores && minmax_dev && pred, "null argument");
    const int64_t total = n_heads * n_tails;
    if (total == 0) return LKG_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    threshold_kernel<<<(int)blocks, 256, 0, stream>>>(scores, ld_scores, n_heads, n_tails, minmax_dev, milestone, pred, ld_pred);
    LKG_LAUNCH_CHECK("threshold_kernel");
    return LKG_OK;
}
