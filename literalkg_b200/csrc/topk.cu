// Per-row top-k and rank over a score matrix (the top-k / rank extension of BASELINE.json; the
// reference has no counterpart -- SURVEY.md fact 7 -- its oracle is torch.topk on calc_score's output).
// Declared order: larger score first; equal scores -> lower column first.  Bit exact w.r.t. that rule.
//
// One CTA per row: 4-pass MSB radix select on the order-preserving uint32 encoding of the score finds
// the k-th largest key, a column-ordered scan collects the winners (ties resolved by column), and a
// bitonic sort in shared memory orders the k survivors.
#include "common.cuh"

namespace lkg {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxK = 1024;

__device__ __forceinline__ uint32_t enc(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (kThreads / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, w, o);
            if (lane >= o) w += t;
        }
        if (lane < (kThreads / 32)) warp_sums[lane] = w;
    }
    __syncthreads();
    const int base = warp ? warp_sums[warp - 1] : 0;
    *total = warp_sums[kThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(kThreads)
topk_rows_kernel(const float* __restrict__ scores, int64_t ld, int64_t n_cols, int k,
                 float* __restrict__ top_val, int64_t* __restrict__ top_col,
                 const int64_t* __restrict__ target, int64_t* __restrict__ ranks) {
    __shared__ int hist[256];
    __shared__ int warp_sums[kThreads / 32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_need, s_cnt;
    __shared__ unsigned long long cand[kMaxK];   // (key << 32) | (~col) so that a descending sort gives the rule

    const int64_t row = blockIdx.x;
    const float* s = scores + row * ld;
    const int tid = threadIdx.x;
    const int kk = (int)min((int64_t)k, n_cols);

    // ---- radix select: find key T of the kk-th largest element --------------------------------
    if (tid == 0) {
        s_prefix = 0;
        s_need = kk;
    }
    __syncthreads();
    for (int pass = 3; pass >= 0 && kk > 0; --pass) {
        hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t mask = pass == 3 ? 0u : (0xffffffffu << (8 * (pass + 1)));
        for (int64_t c = tid; c < n_cols; c += kThreads) {
            const uint32_t key = enc(s[c]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int need = s_need, b = 255;
            for (; b > 0; --b) {
                if (hist[b] >= need) break;
                need -= hist[b];
            }
            s_need = need;                       // how many we still need inside bucket b
            s_prefix = prefix | ((uint32_t)b << (8 * pass));
        }
        __syncthreads();
    }
    const uint32_t T = s_prefix;
    const int need_eq = s_need;                  // elements == T to take, lowest column first

    // ---- collect: all keys > T, and the first need_eq keys == T in column order ----------------
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int eq_taken = 0;
    for (int64_t c0 = 0; c0 < n_cols && kk > 0; c0 += kThreads) {
        const int64_t c = c0 + tid;
        uint32_t key = 0;
        bool gt = false, eq = false;
        if (c < n_cols) {
            key = enc(s[c]);
            gt = key > T;
            eq = key == T;
        }
        int total_eq;
        const int eq_rank = block_exclusive_scan(eq ? 1 : 0, warp_sums, &total_eq);
        const bool take = gt || (eq && eq_taken + eq_rank < need_eq);
        if (take) {
            const int slot = atomicAdd(&s_cnt, 1);
            cand[slot] = ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)c);
        }
        eq_taken += total_eq;
    }
    __syncthreads();

    // ---- bitonic sort (descending) of the kk candidates, padded to a power of two --------------
    int p2 = 1;
    while (p2 < kk) p2 <<= 1;
    for (int i = kk + tid; i < p2; i += kThreads) cand[i] = 0ull;
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < p2; i += kThreads) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const unsigned long long a = cand[i], b = cand[j];
                    if ((a < b) == desc) {
                        cand[i] = b;
                        cand[j] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += kThreads) {
        if (i < kk) {
            top_val[row * k + i] = dec((uint32_t)(cand[i] >> 32));
            top_col[row * k + i] = (int64_t)(~(uint32_t)(cand[i] & 0xffffffffu));
        } else {
            top_val[row * k + i] = -INFINITY;
            top_col[row * k + i] = -1;
        }
    }

    // ---- rank of a target column -----------------------------------------------------------------
    if (target && ranks) {
        const int64_t tc = target[row];
        const uint32_t tkey = enc(s[tc]);
        int better = 0;
        for (int64_t c = tid; c < n_cols; c += kThreads) {
            const uint32_t key = enc(s[c]);
            better += (key > tkey || (key == tkey && c < tc)) ? 1 : 0;
        }
        int total;
        block_exclusive_scan(better, warp_sums, &total);
        if (tid == 0) ranks[row] = total;
    }
}

// ---- k-way merge of per-rank survivor lists (sharded scoring) ---------------------------------------
// vals / ids [parts][n_rows][k]: every rank's k best (score, global tail position) per head, -inf / -1 padded.  One warp
// per head sorts the parts * k <= 1024 keys (score desc, position asc) in shared memory and writes the first k.
constexpr int kMergeMax = 1024;
__global__ void __launch_bounds__(128) topk_merge_kernel(const float* __restrict__ vals, const int64_t* __restrict__ ids,
                                                         int parts, int64_t n_rows, int k, float* __restrict__ top_val,
                                                         int64_t* __restrict__ top_id) {
    __shared__ unsigned long long keys[kMergeMax];
    const int64_t row = blockIdx.x;
    const int w = parts * k;
    int p2 = 1;
    while (p2 < w) p2 <<= 1;
    for (int i = threadIdx.x; i < p2; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < w) {
            const int part = i / k, j = i - part * k;
            const int64_t src = ((int64_t)part * n_rows + row) * k + j;
            const int64_t id = ids[src];
            if (id >= 0) key = ((unsigned long long)enc(vals[src]) << 32) | (uint32_t)(~(uint32_t)id);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const unsigned long long a = keys[i], b = keys[j];
                    if ((a < b) == desc) {
                        keys[i] = b;
                        keys[j] = a;
                    }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const unsigned long long key = i < p2 ? keys[i] : 0ull;
        const bool ok = key != 0ull;
        top_val[row * k + i] = ok ? dec((uint32_t)(key >> 32)) : -INFINITY;
        top_id[row * k + i] = ok ? (int64_t)(~(uint32_t)(key & 0xffffffffu)) : -1;
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_topk_merge(const float* vals, const int64_t* ids, int32_t parts, int64_t n_rows, int32_t k,
                              float* top_values, int64_t* top_ids, void* stream_) {
    LKG_REQUIRE(vals && ids && top_values && top_ids && parts >= 1 && k >= 1 && n_rows >= 0, "bad merge arguments");
    LKG_REQUIRE((int64_t)parts * k <= kMergeMax, "parts * k must be <= %d (got %d x %d)", kMergeMax, parts, k);
    if (n_rows == 0) return LKG_OK;
    topk_merge_kernel<<<(unsigned)n_rows, 128, 0, (cudaStream_t)stream_>>>(vals, ids, parts, n_rows, k, top_values, top_ids);
    LKG_LAUNCH_CHECK("topk_merge_kernel");
    return LKG_OK;
}

extern "C" int lkg_topk_rows(const float* scores, int64_t ld_scores, int64_t n_rows, int64_t n_cols, int32_t k,
                             float* top_values, int64_t* top_cols, const int64_t* target_cols,
                             int64_t* ranks_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(scores && top_values && top_cols, "null argument");
    LKG_REQUIRE(k >= 1 && k <= kMaxK, "k must be in [1, %d] (got %d)", kMaxK, k);
    LKG_REQUIRE(n_cols >= 1 && n_cols < (1ll << 32) && ld_scores >= n_cols, "bad column count");
    if (n_rows == 0) return LKG_OK;
    topk_rows_kernel<<<(unsigned)n_rows, kThreads, 0, stream>>>(scores, ld_scores, n_cols, k, top_values, top_cols,
                                                               target_cols, ranks_out);
    LKG_LAUNCH_CHECK("topk_rows_kernel");
    return LKG_OK;
}
