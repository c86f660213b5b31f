// Parameter-gradient GEMM of the backward pass on the tensor cores:  C[dx, cy] += X^T Y  reduced over the N entity rows
// (d W = d out^T @ input of every Linear the embedding pass contains: gate.py:22-28, model.py:90-130, 309-310).
//
// The reduction runs over the ROWS of two row-major matrices, i.e. both MMA operands are MN-major: the same fp16
// hi/lo planes the forward GEMMs read ([rows, cols], cols contiguous) are staged by TMA as 64-row x 64-column boxes
// (128-byte swizzle; one box row = 64 consecutive columns of one entity row) and described to tcgen05.mma with
// MN-major shared-memory descriptors (canonical layout ((8,m),(8,k)) : ((1,LBO),(8,SBO)) in 16-byte units:
// SBO = 1024 B between 8-row groups, LBO = the distance between 64-column blocks) -- no transposed copy of either
// operand is ever written.  Three products per k-step (hi*hi + lo*hi + hi*lo) accumulate in fp32 TMEM, like the
// forward engine (gemm_tc.cu).
//
// Split-K: tile (m, n) x row range per CTA, the partial tiles are added to the fp32 output with atomics
// (the output is at most 600 x 602).  HBM bound: every X column block is read tiles_n times, every Y block tiles_m
// times.
#include "tc.cuh"

namespace lkg {
namespace {

using namespace tc;

constexpr int kRowsPerChunk = 64;                       // K extent of one pipeline stage
constexpr int kBlockCols = 64;                          // columns per TMA box == one 128-byte swizzle row
constexpr uint32_t kBlockBytes = 2 * kRowsPerChunk * kBlockCols * 2;   // hi + lo planes of one box: 16 KB
constexpr int kXBlocks = kBM / kBlockCols;              // 2
constexpr int kMaxYBlocks = 4;                          // N tile <= 256
constexpr int kXtyStages = 2;
constexpr uint32_t kXtyStageBytes = (kXBlocks + kMaxYBlocks) * kBlockBytes;   // 96 KB
constexpr uint32_t kXtySmem = kXtyStages * kXtyStageBytes + 1024 + 256;
constexpr int kXtyThreads = 192;

struct XtyTcParams {
    CUtensorMap x_map, y_map;
    int64_t n_rows;
    int dx, cy, bn, y_blocks;
    int tiles_m, tiles_n, splits;
    int64_t rows_per_split;      // multiple of kRowsPerChunk
    const float* x_rec;
    const float* y_rec;
    float* out;
    int64_t ld_out;
};

// MN-major, 128-byte swizzle: start >> 4, LBO (bytes >> 4) between 64-element MN blocks, SBO = 8 rows x 128 B
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16, D = f32, A = B = fp16, both MN-major (bits 15 / 16), M = 128, N = bn
__device__ __forceinline__ uint32_t umma_idesc_mn(int bn) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

__global__ void __launch_bounds__(kXtyThreads, 1) tc_xty_kernel(const __grid_constant__ XtyTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kXtyStages * kXtyStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kXtyStages;
    uint64_t* acc_full = bars + 2 * kXtyStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kXtyStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int tile = blockIdx.x;
    const int ks = tile % p.splits;
    tile /= p.splits;
    const int nb = tile % p.tiles_n;
    const int mb = tile / p.tiles_n;
    const int64_t r_begin = (int64_t)ks * p.rows_per_split;
    const int64_t r_end = min(p.n_rows, r_begin + p.rows_per_split);
    const int n_chunks = r_end > r_begin ? (int)((r_end - r_begin + kRowsPerChunk - 1) / kRowsPerChunk) : 0;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.x_map) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.y_map) : "memory");
        for (int s = 0; s < kXtyStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t stage_tx = (uint32_t)(kXBlocks + p.y_blocks) * kBlockBytes;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + stage * kXtyStageBytes;
                mbar_expect_tx(&full[stage], stage_tx);
                const int row = (int)(r_begin + (int64_t)c * kRowsPerChunk);
                for (int b = 0; b < kXBlocks; ++b)
                    tma_load_3d(&p.x_map, &full[stage], st + b * kBlockBytes, mb * kBM + b * kBlockCols, row, 0);
                for (int b = 0; b < p.y_blocks; ++b)
                    tma_load_3d(&p.y_map, &full[stage], st + (kXBlocks + b) * kBlockBytes, nb * p.bn + b * kBlockCols,
                                row, 0);
                if (++stage == kXtyStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_mn(p.bn);
            uint32_t stage = 0, phase = 0;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t x_hi = smem_u32(smem + stage * kXtyStageBytes);
                const uint32_t x_lo = x_hi + kBlockBytes / 2;
                const uint32_t y_hi = x_hi + kXBlocks * kBlockBytes;
                const uint32_t y_lo = y_hi + kBlockBytes / 2;
#pragma unroll
                for (int k = 0; k < kRowsPerChunk / 16; ++k) {
                    const uint32_t koff = k * 16 * 128;            // 16 rows of 128 bytes
                    const uint64_t dxh = umma_desc_mn(x_hi + koff, kBlockBytes), dxl = umma_desc_mn(x_lo + koff, kBlockBytes);
                    const uint64_t dyh = umma_desc_mn(y_hi + koff, kBlockBytes), dyl = umma_desc_mn(y_lo + koff, kBlockBytes);
                    tc_mma_f16(tmem_base, dxh, dyh, idesc, (c | k) != 0);
                    tc_mma_f16(tmem_base, dxl, dyh, idesc, 1);
                    tc_mma_f16(tmem_base, dxh, dyl, idesc, 1);
                }
                tc_commit(&empty[stage]);
                if (++stage == kXtyStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            tc_commit(acc_full);
        }
    } else if (n_chunks > 0) {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const float scale = __ldg(p.x_rec + 2) * __ldg(p.y_rec + 2);       // powers of two: exact
        const int i = mb * kBM + warp * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < p.bn; c0 += 16) {
            float v[16];
            tc_ld16(taddr + c0, v);
            if (i < p.dx) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = nb * p.bn + c0 + j;
                    if (col < p.cy) atomicAdd(p.out + (int64_t)i * p.ld_out + col, v[j] * scale);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

int make_rows_map(CUtensorMap* map, const void* base, int64_t rows, int k, int64_t ld, int64_t plane_stride) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) LKG_FAIL(LKG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    LKG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (ld * 2) % 16 == 0 && (plane_stride * 2) % 16 == 0,
                "fp16 planes must be 16-byte aligned (base, row stride, plane stride)");
    cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
    cuuint32_t box[3] = {(cuuint32_t)kBlockCols, (cuuint32_t)kRowsPerChunk, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LKG_FAIL(LKG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LKG_OK;
}

// column sums: out[j] += sum_rows y[row, j]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ y, int64_t ld, int64_t n, int c,
                                                     int64_t rows_per_cta, float* __restrict__ out) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = min(n, r0 + rows_per_cta);
    for (int j = threadIdx.x; j < c; j += blockDim.x) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int64_t r = r0;
        for (; r + 3 < r1; r += 4) {
            s0 += __ldg(y + r * ld + j);
            s1 += __ldg(y + (r + 1) * ld + j);
            s2 += __ldg(y + (r + 2) * ld + j);
            s3 += __ldg(y + (r + 3) * ld + j);
        }
        for (; r < r1; ++r) s0 += __ldg(y + r * ld + j);
        atomicAdd(out + j, (s0 + s1) + (s2 + s3));
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_xt_y_planes(const lkg_planes* x, const lkg_planes* y, int64_t n_rows, float* out, int64_t ld_out,
                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(x && y && out && x->n_segments == 1 && y->n_segments == 1, "xt_y takes single-segment planes");
    LKG_REQUIRE(x->scale[0] && y->scale[0], "both operands need scale records");
    if (n_rows == 0) return LKG_OK;
    XtyTcParams p{};
    p.n_rows = n_rows;
    p.dx = x->k[0];
    p.cy = y->k[0];
    LKG_REQUIRE(p.dx > 0 && p.cy > 0 && ld_out >= p.cy, "bad xt_y shape");
    const int y_cols = (p.cy + kBlockCols - 1) / kBlockCols;                  // 64-column blocks of Y
    p.tiles_n = (y_cols + kMaxYBlocks - 1) / kMaxYBlocks;
    p.y_blocks = (y_cols + p.tiles_n - 1) / p.tiles_n;
    p.bn = p.y_blocks * kBlockCols;
    p.tiles_m = (p.dx + kBM - 1) / kBM;
    const int tiles = p.tiles_m * p.tiles_n;
    const int64_t chunks = (n_rows + kRowsPerChunk - 1) / kRowsPerChunk;
    int64_t splits = (2 * (int64_t)sm_count() + tiles - 1) / tiles;
    if (splits > chunks) splits = chunks;
    if (splits < 1) splits = 1;
    p.rows_per_split = (chunks + splits - 1) / splits * kRowsPerChunk;
    p.splits = (int)((n_rows + p.rows_per_split - 1) / p.rows_per_split);
    p.x_rec = x->scale[0];
    p.y_rec = y->scale[0];
    p.out = out;
    p.ld_out = ld_out;
    if (int rc = make_rows_map(&p.x_map, x->ptr[0], n_rows, p.dx, x->ld[0], x->plane_stride[0])) return rc;
    if (int rc = make_rows_map(&p.y_map, y->ptr[0], n_rows, p.cy, y->ld[0], y->plane_stride[0])) return rc;
    LKG_CUDA(cudaFuncSetAttribute(tc_xty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXtySmem));
    tc_xty_kernel<<<tiles * p.splits, kXtyThreads, kXtySmem, stream>>>(p);
    LKG_LAUNCH_CHECK("tc_xty_kernel");
    return LKG_OK;
}

extern "C" int lkg_colsum(const float* y, int64_t ld_y, int64_t n, int32_t c, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0 || c == 0) return LKG_OK;
    LKG_REQUIRE(y && out && c > 0, "bad colsum arguments");
    int64_t ctas = (int64_t)sm_count() * 8;
    int64_t rows_per_cta = (n + ctas - 1) / ctas;
    if (rows_per_cta < 16) rows_per_cta = 16;
    ctas = (n + rows_per_cta - 1) / rows_per_cta;
    colsum_kernel<<<(unsigned)ctas, 256, 0, stream>>>(y, ld_y, n, c, rows_per_cta, out);
    LKG_LAUNCH_CHECK("colsum_kernel");
    return LKG_OK;
}
