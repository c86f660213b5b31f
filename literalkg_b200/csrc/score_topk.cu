// All-entity link-prediction scoring with a fused top-k epilogue (BASELINE.json north star (d); extension of
// calc_score, model.py:473-486 -- the reference has no top-k; its oracle is torch.topk applied to
// calc_score's output).  The B_h x N_t score matrix is never materialised.
//
//   lkg_score_index      fp32 embedding rows -> scaled fp16 "hi" plane + row norms + max norm (one pass)
//   lkg_score_topk       1. threshold: thr_h = theta_h - E_h, theta_h = k-th best EXACT score of the head among a
//                           sample of the tails (so the true k-th best is >= theta_h), E_h = rigorous bound of the
//                           single-product fp16 error, c * |h| * max_t |t|;
//                        2. filter GEMM (tcgen05, ONE fp16 product per k-step): every (head, tail) whose approximate
//                           score reaches thr_h is appended to the head's candidate list -- the true top-k is a
//                           subset by construction;
//                        3. finalize: candidates re-scored exactly (fp32 products summed in fp64, rounded once to
//                           fp32), sorted by (score desc, column asc), first k written out.  A head whose list
//                           overflowed falls back to an exact scan of all tails on the device (slow, correct).
//
// sm_100a design of the filter GEMM
//   * persistent CTA per SM; the A operand (two 128-row head blocks x K <= 256, <= 128 KB) is loaded by TMA once and
//     stays resident in shared memory; tail tiles (128 rows x 64-wide K chunks, 16 KB) stream through a 5-stage
//     mbarrier ring -- per tile one B fill feeds two MMAs, which halves the L2->SM traffic per flop (a 128 x 256
//     tile with both operands streamed needs ~175 GB/s per SM, more than L2 can deliver to 148 SMs);
//   * CTAs that share a tail tile (different head blocks) have consecutive ids, so a tile is read from HBM once
//     and hit in L2 by the other head blocks;
//   * accumulators: 2 buffers x 2 head blocks x 128 TMEM columns (all 512), epilogue of tile i overlaps MMAs of i+1;
//   * epilogue: 4 warps, thread = head row, tcgen05.ld 32 columns at a time, max-reduce, ONE compare against the
//     row's threshold (pre-multiplied by the operand scales).  Candidates go to a list private to this (head row,
//     CTA stream): exactly one thread of one CTA writes it, so the counter lives in a register and there is no
//     atomic.  They are staged in shared memory, 16 per row, and flushed to the global list when the stage fills
//     up: a global store (let alone an atomicAdd round trip) in front of the releasing mbarrier.arrive that hands
//     the accumulator back made the epilogue wait for the store to land on most tiles (measured 3.6 ms vs 1.1 ms).
#include "tc.cuh"

namespace lkg {
namespace {

using namespace tc;

constexpr int kBN = 128;                 // tails per tile
constexpr int kStages = 5;
constexpr int kThreads = 192;            // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
constexpr int kMaxChunks = 4;            // K <= 256
constexpr uint32_t kChunkBytes = kBM * kBK * 2;     // 16 KB: 128 rows x 64 fp16 (A block chunk and B stage alike)
constexpr int kStageCand = 16;                      // candidates staged in shared memory per (thread, head block)
constexpr float kErrCoef = 1.15f / 1024.f;          // |s~ - s| <= kErrCoef |h| |t|: two fp16 roundings (2 * 2^-11),
                                                    // fp32 accumulation over K <= 256 (2^-16), theta's own 2^-21
constexpr int kFinalThreads = 256;

struct FilterParams {
    CUtensorMap a_map;       // heads hi plane [n_heads, K] fp16
    CUtensorMap b_map;       // tails hi plane [n_tails, K] fp16
    int n_heads, n_tails, n_chunks, n_pairs, n_tiles;
    const float* thr;        // [n_heads] threshold in accumulator (scaled) units
    int* cnt;                // [n_heads, n_streams] candidates seen by each CTA stream
    int* cand;               // [n_heads, n_streams, cap_s] candidate columns (positions in the tail list)
    int cap_s;               // slots per (head, stream) sublist
    // sampling mode (tilemax != NULL): visit tiles 0, tile_stride, 2 tile_stride, ... (n_tiles of them) and store
    // every head's largest approximate score of each visited tile instead of filtering
    float* tilemax;          // [n_heads, n_tiles]
    int tile_stride;
};

__global__ void __launch_bounds__(kThreads, 1) score_filter_kernel(const __grid_constant__ FilterParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                           // [2 blocks][n_chunks] x 16 KB
    uint8_t* smem_b = smem + 2 * kMaxChunks * kChunkBytes;            // [kStages] x 16 KB
    int* smem_cand = reinterpret_cast<int*>(smem_b + kStages * kChunkBytes);        // [2][kStageCand][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_cand + 2 * kStageCand * 128);
    uint64_t* full = bars;                    // [kStages]
    uint64_t* empty = bars + kStages;         // [kStages]
    uint64_t* acc_full = bars + 2 * kStages;  // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint64_t* a_full = acc_empty + 2;         // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.x % p.n_pairs;
    const int stream = blockIdx.x / p.n_pairs, n_streams = gridDim.x / p.n_pairs;
    const int row_base = pair * 2 * kBM;
    const int n_blk = row_base + kBM < p.n_heads ? 2 : 1;             // a pair whose second block is empty skips it

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.a_map) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&p.b_map) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 128);
        }
        mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(a_full, (uint32_t)(n_blk * p.n_chunks) * kChunkBytes);
            for (int b = 0; b < n_blk; ++b)
                for (int kc = 0; kc < p.n_chunks; ++kc)
                    tma_load_2d(&p.a_map, a_full, smem_a + (b * kMaxChunks + kc) * kChunkBytes, kc * kBK,
                                row_base + b * kBM);
            uint32_t stage = 0, phase = 0;
            for (int tile = stream; tile < p.n_tiles; tile += n_streams) {
                for (int kc = 0; kc < p.n_chunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], kChunkBytes);
                    tma_load_2d(&p.b_map, &full[stage], smem_b + stage * kChunkBytes, kc * kBK,
                                tile * p.tile_stride * kBN);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(kBN);
            mbar_wait(a_full, 0);
            tc_fence_after();
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = stream; tile < p.n_tiles; tile += n_streams, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kc = 0; kc < p.n_chunks; ++kc) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(smem_b + stage * kChunkBytes);
                    for (int b = 0; b < n_blk; ++b) {
                        const uint32_t a_addr = smem_u32(smem_a + (b * kMaxChunks + kc) * kChunkBytes);
                        const uint32_t tmem_d = tmem_base + buf * 2 * kBN + b * kBN;
#pragma unroll
                        for (int k = 0; k < kBK / 16; ++k)
                            tc_mma_f16(tmem_d, umma_desc(a_addr + k * 32), umma_desc(b_addr + k * 32), idesc,
                                       (kc | k) != 0);
                    }
                    tc_commit(&empty[stage]);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&acc_full[buf]);
            }
        }
    } else {
        // ===== epilogue warps 0-3: TMEM lane = head row of the block =====
        // The candidate path is rare per thread but hit by some lane of the warp on most tiles, so it is kept small
        // (a bit mask + a compact loop, no unrolled copies): a 32-way unrolled version made the kernel 600 KB of
        // SASS and every entry into it an instruction-cache miss (3.6 ms instead of 1.1 ms).
        int row[2], seen[2] = {0, 0}, staged[2] = {0, 0};
        float thr[2];
        int* my_cand[2];
        int* my_stage = smem_cand + threadIdx.x;        // slot i of block b: my_stage[(b * kStageCand + i) * 128]
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            row[b] = row_base + b * kBM + warp * 32 + lane;
            thr[b] = (row[b] < p.n_heads && !p.tilemax) ? __ldg(p.thr + row[b]) : INFINITY;
            my_cand[b] = p.cand + ((int64_t)row[b] * n_streams + stream) * p.cap_s;
        }
        int it = 0;
        for (int tile = stream; tile < p.n_tiles; tile += n_streams, ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (it >> 1) & 1);
            tc_fence_after();
            const int tile_col0 = tile * p.tile_stride * kBN;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                if (b >= n_blk) break;
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 2 * kBN + b * kBN;
                float tmax = -INFINITY;
#pragma unroll 1
                for (int c0 = 0; c0 < kBN; c0 += 32) {
                    uint32_t r[32];
                    tc_ld32_nowait(taddr + c0, r);
                    tc_ld_wait();
                    const int col0 = tile_col0 + c0;
                    const int valid = p.n_tails - col0;                  // columns of this group inside the tail list
                    float m = __uint_as_float(r[0]);
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
                    if (p.tilemax) {
                        if (valid < 32) {                                // last, partial tile: TMA zero-filled the rest
                            m = -INFINITY;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < valid) m = fmaxf(m, __uint_as_float(r[j]));
                        }
                        tmax = fmaxf(tmax, m);
                    } else if (m >= thr[b]) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(r[j]) >= thr[b] ? 1u : 0u) << j;
                        if (valid < 32) mask &= valid <= 0 ? 0u : (1u << valid) - 1u;
#pragma unroll 1
                        while (mask) {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            my_stage[(b * kStageCand + staged[b]) * 128] = col0 + j;
                            if (++staged[b] == kStageCand) {             // stage full: flush to the global list
#pragma unroll 1
                                for (int i = 0; i < kStageCand; ++i)
                                    if (seen[b] + i < p.cap_s) my_cand[b][seen[b] + i] = my_stage[(b * kStageCand + i) * 128];
                                seen[b] += kStageCand;
                                staged[b] = 0;
                            }
                        }
                    }
                }
                if (p.tilemax && row[b] < p.n_heads) p.tilemax[(int64_t)row[b] * p.n_tiles + tile] = tmax;
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
        if (!p.tilemax) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll 1
                for (int i = 0; i < staged[b]; ++i)
                    if (seen[b] + i < p.cap_s) my_cand[b][seen[b] + i] = my_stage[(b * kStageCand + i) * 128];
                seen[b] += staged[b];
                if (row[b] < p.n_heads) p.cnt[(int64_t)row[b] * n_streams + stream] = seen[b];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

constexpr uint32_t kFilterSmem = (2 * kMaxChunks + kStages) * kChunkBytes + 2 * kStageCand * 128 * 4 /*candidates*/ +
                                 1024 /*align*/ + 256 /*barriers*/;
static_assert(kFilterSmem <= 227 * 1024, "filter kernel shared memory");

int make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int k, int64_t ld) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) LKG_FAIL(LKG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    LKG_REQUIRE(aligned16(base) && (ld * 2) % 16 == 0, "hi planes must be 16-byte aligned (base, row stride)");
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) LKG_FAIL(LKG_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LKG_OK;
}

// ---- index: fp32 rows -> scaled fp16 hi plane + norms ------------------------------------------------------------
__global__ void score_index_kernel(const float* __restrict__ emb, int64_t ld, const int64_t* __restrict__ rows,
                                   int64_t m, int dim, const float* __restrict__ rec, __half* __restrict__ hi,
                                   int64_t ld_hi, float* __restrict__ norms, float* __restrict__ max_norm) {
    const float scale = __ldg(rec + 1);
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int groups = (int)(ld_hi >> 3);
    float wmax = 0.f;
    for (int64_t r = warp; r < m; r += nwarps) {
        const float* src = emb + (rows ? rows[r] : r) * ld;
        float ss = 0.f;
        for (int gq = lane; gq < groups; gq += 32) {
            const int c0 = 8 * gq;
            float x[8];
            if (c0 + 8 <= dim && (ld & 3) == 0) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(src + c0));
                const float4 b = __ldg(reinterpret_cast<const float4*>(src + c0) + 1);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = c0 + j < dim ? __ldg(src + c0 + j) : 0.f;
            }
            __align__(16) __half h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = x[j] * scale;
                h[j] = __float2half_rn(v);
                ss = fmaf(v, v, ss);
            }
            *reinterpret_cast<uint4*>(hi + r * ld_hi + c0) = *reinterpret_cast<const uint4*>(h);
        }
        const float nrm = sqrtf(warp_sum(ss)) * 1.000001f;     // scaled units, rounded up
        if (lane == 0) norms[r] = nrm;
        wmax = fmaxf(wmax, nrm);
    }
    if (lane == 0 && wmax > 0.f) atomicMax(reinterpret_cast<uint32_t*>(max_norm), __float_as_uint(wmax));
}

// E (accumulator units): bound of |single-product fp16 score - exact score| for one head against any tail
__device__ __forceinline__ float score_err(float nh, float nt, int dim) {
    return kErrCoef * nh * nt + 4.8e-7f /*2^-21*/ * (nh + nt) * sqrtf((float)dim) + 1.f;
}

// thr = theta * scale^2 - E with theta a lower bound of the head's k-th best EXACT score (caller supplied)
__global__ void score_threshold_kernel(const float* __restrict__ theta, int64_t theta_stride, int n_heads,
                                       const float* __restrict__ head_norms, const float* __restrict__ tail_max_norm,
                                       const float* __restrict__ rec, int dim, float* __restrict__ thr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_heads) return;
    const float s2 = rec[1] * rec[1];
    const float err = score_err(head_norms[i], *tail_max_norm, dim);
    const float th = theta ? theta[(int64_t)i * theta_stride] : -INFINITY;
    // theta == -inf (no bound known): every tail is a candidate
    thr[i] = th == -INFINITY ? -INFINITY : th * s2 - err - fabsf(th * s2) * 1e-6f;
}

// thr = (k-th largest tile maximum of the sampled tiles) - 2E.  The tile maxima are k distinct tails whose
// APPROXIMATE scores reach m_k, so their exact scores reach m_k - E, so the exact k-th best is >= m_k - E and every
// member of the true top-k has an approximate score >= m_k - 2E.  One CTA per head, bitonic sort in shared memory.
__global__ void __launch_bounds__(256) score_sample_threshold_kernel(const float* __restrict__ tilemax, int n_st, int k,
                                                                     const float* __restrict__ head_norms,
                                                                     const float* __restrict__ tail_max_norm, int dim,
                                                                     float* __restrict__ thr) {
    extern __shared__ float sv[];
    const int head = blockIdx.x;
    int p2 = 1;
    while (p2 < n_st) p2 <<= 1;
    for (int i = threadIdx.x; i < p2; i += blockDim.x) sv[i] = i < n_st ? tilemax[(int64_t)head * n_st + i] : -INFINITY;
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const float a = sv[i], b = sv[j];
                    if ((a < b) == desc) {
                        sv[i] = b;
                        sv[j] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        const float mk = k <= n_st ? sv[k - 1] : -INFINITY;
        thr[head] = mk == -INFINITY ? -INFINITY : mk - 2.f * score_err(head_norms[head], *tail_max_norm, dim) - fabsf(mk) * 1e-6f;
    }
}

// ---- finalize ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t enc(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Exact score of one (head, tail): fp32 products are exact in fp64, summed in fp64, one rounding to fp32 at the end.
// Eight lanes share a tail row (lane q of the group takes float4 q, q + 8, ...), so a warp scores four candidates
// at once with all its row loads in flight together; the head row is read from shared memory.
__device__ __forceinline__ float exact_dot8(const float* __restrict__ s_head, const float* __restrict__ trow, int dim,
                                            int q, bool live) {
    double acc = 0.0;
    const int nv = dim >> 2;                                   // dim % 4 == 0 is checked on the host side
    float4 t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int v = q + 8 * i;
        t[i] = (live && v < nv) ? __ldg(reinterpret_cast<const float4*>(trow) + v) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int v = q + 8 * i;
        if (v < nv) {
            const float4 h = *reinterpret_cast<const float4*>(s_head + 4 * v);
            acc = fma((double)h.x, (double)t[i].x, acc);
            acc = fma((double)h.y, (double)t[i].y, acc);
            acc = fma((double)h.z, (double)t[i].z, acc);
            acc = fma((double)h.w, (double)t[i].w, acc);
        }
    }
    acc += __shfl_xor_sync(kFull, acc, 4);
    acc += __shfl_xor_sync(kFull, acc, 2);
    acc += __shfl_xor_sync(kFull, acc, 1);
    return (float)acc;
}

__device__ void bitonic_desc(unsigned long long* keys, int n_pow2) {
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
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
    }
}

struct FinalParams {
    const float* emb;           // tails' matrix
    int64_t ld_emb;
    const float* head_emb;      // heads' matrix (may be the same)
    int64_t ld_head_emb;
    const int64_t* head_rows;   // nullable: head i = row head_row_base + i
    int64_t head_row_base;
    const int64_t* tail_rows;   // nullable: tail j = row j
    int n_heads, n_tails, dim, k, cap;
    int n_streams, cap_s;       // candidate sublists per head and their capacity (n_streams * cap_s <= cap)
    const int* cnt;
    const int* cand;
    float* top_val;
    int64_t* top_col;
};

// One CTA per head.  keys: (ordered score << 32) | ~column, so a descending sort gives "larger score first, ties ->
// lower column".
__global__ void __launch_bounds__(kFinalThreads) score_finalize_kernel(FinalParams p) {
    extern __shared__ unsigned long long keys[];
    __shared__ int s_kept;
    const int head = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = kFinalThreads / 32;
    const int grp = lane >> 3, q = lane & 7;                    // 4 candidates per warp, 8 lanes each
    const int kk = min(p.k, p.n_tails);
    __shared__ __align__(16) float s_head[kMaxChunks * kBK];
    const float* hrow = p.head_emb + (p.head_rows ? p.head_rows[head] : p.head_row_base + head) * p.ld_head_emb;
    for (int i = threadIdx.x; i < kMaxChunks * kBK; i += blockDim.x) s_head[i] = i < p.dim ? __ldg(hrow + i) : 0.f;
    if (threadIdx.x == 0) s_kept = 0;
    // gather the head's sublists (one per CTA stream of the filter kernel) into one column list
    int* cols = reinterpret_cast<int*>(keys + p.cap);
    __shared__ int s_off[161], s_n, s_over;
    for (int sidx = threadIdx.x; sidx < p.n_streams; sidx += blockDim.x)
        s_off[sidx] = p.cnt[(int64_t)head * p.n_streams + sidx];
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0, over = 0;
        for (int sidx = 0; sidx < p.n_streams; ++sidx) {
            const int c = s_off[sidx];
            over |= c > p.cap_s;
            s_off[sidx] = tot;
            tot += min(c, p.cap_s);
        }
        s_off[p.n_streams] = tot;
        s_n = tot;
        s_over = over;
    }
    __syncthreads();
    const bool overflow = s_over != 0;
    int n = 0;
    auto tail_row = [&](int col) { return p.emb + (p.tail_rows ? p.tail_rows[col] : col) * p.ld_emb; };
    if (!overflow) {
        n = s_n;
        for (int sidx = warp; sidx < p.n_streams; sidx += nwarps) {
            const int o = s_off[sidx], c = s_off[sidx + 1] - o;
            const int* src = p.cand + ((int64_t)head * p.n_streams + sidx) * p.cap_s;
            for (int i = lane; i < c; i += 32) cols[o + i] = src[i];
        }
        __syncthreads();
        for (int c0 = 4 * warp; c0 < n; c0 += 4 * nwarps) {
            const int c = c0 + grp;
            const bool live = c < n;
            const int col = live ? cols[c] : 0;
            const float s = exact_dot8(s_head, tail_row(col), p.dim, q, live);
            if (live && q == 0) keys[c] = ((unsigned long long)enc(s) << 32) | (uint32_t)(~(uint32_t)col);
        }
    } else {
        // Overflow fallback (the candidate band held more than `cap` tails, e.g. a plateau of near-identical
        // scores): exact scan of every tail, keeping the best `cap / 2` keys seen so far: when the buffer fills up
        // it is sorted and its lower half dropped (cap / 2 >= k is checked on the host side).
        const int half = p.cap / 2;
        for (int base = 0; base < p.n_tails; base += half) {
            const int chunk = min(half, p.n_tails - base);
            const int kept = s_kept;
            for (int c0 = 4 * warp; c0 < chunk; c0 += 4 * nwarps) {
                const int c = c0 + grp;
                const bool live = c < chunk;
                const int col = live ? base + c : 0;
                const float s = exact_dot8(s_head, tail_row(col), p.dim, q, live);
                if (live && q == 0) keys[kept + c] = ((unsigned long long)enc(s) << 32) | (uint32_t)(~(uint32_t)col);
            }
            __syncthreads();
            const int tot = kept + chunk;
            int p2 = 1;
            while (p2 < tot) p2 <<= 1;
            for (int i = tot + threadIdx.x; i < p2; i += blockDim.x) keys[i] = 0ull;
            __syncthreads();
            bitonic_desc(keys, p2);
            if (threadIdx.x == 0) s_kept = min(tot, half);
            __syncthreads();
        }
        n = s_kept;
    }
    __syncthreads();
    if (!overflow) {
        int p2 = 1;
        while (p2 < n) p2 <<= 1;
        for (int i = n + threadIdx.x; i < p2; i += blockDim.x) keys[i] = 0ull;
        __syncthreads();
        bitonic_desc(keys, p2);
    }
    for (int i = threadIdx.x; i < p.k; i += blockDim.x) {
        const bool ok = i < kk && i < n;
        p.top_val[(int64_t)head * p.k + i] = ok ? dec((uint32_t)(keys[i] >> 32)) : -INFINITY;
        p.top_col[(int64_t)head * p.k + i] = ok ? (int64_t)(~(uint32_t)(keys[i] & 0xffffffffu)) : -1;
    }
}

// ---- rank of a given tail per head (north star (d): "top-k and rank epilogue") -----------------------------------
// rank_i = #{ j : s_ij > s_it  or  (s_ij == s_it and j < t_i) } with s the EXACT scores (fp32 products summed in fp64,
// one rounding -- the values the fused top-k returns) and t_i the position of head i's target in the tail list.
//   1. lkg_rank_prepare   tau_i = exact score of (head i, target i); band tau_i -/+ E_i with E_i a bound of the
//                         3-product fp16 hi/lo GEMM's error for K <= 256: representation (operands kept to 2^-22, the
//                         lo * lo product dropped) 3 * 2^-22 |h||t| = 0.7e-6, fp32 accumulation of 48 MMAs at <= 2^-23
//                         of the running magnitude each = 5.7e-6; kRankErr = 1e-5 leaves a factor 1.5 on top;
//   2. lkg_score_rank     (gemm_tc.cu) the score GEMM with a counting epilogue: columns certainly above tau are counted,
//                         the ones inside the band are listed; no score leaves the SM;
//   3. lkg_rank_finalize  the listed columns re-scored exactly and compared with tau under the tie rule.
constexpr float kRankErr = 1.0e-5f;

// center (nullable, [dim]): the GEMM's tails are t_j - center (a common shift moves every score of a head by the same
// h . center, so the ranking is unchanged while the error bound shrinks from |h| max|t| to |h| max|t - center| -- the
// embeddings of a trained or freshly initialised model are tightly clustered around their mean, and a band relative
// to |t| would hold tens of thousands of tails); the thresholds are then taken around tau - h . center.
// out[r, :] = src[rows ? rows[r] : r, :] - center  (the shifted tails of the rank GEMM)
__global__ void shift_rows_kernel(const float* __restrict__ src, int64_t ld, const int64_t* __restrict__ rows, int64_t m,
                                  int k, const float* __restrict__ center, float* __restrict__ out, int64_t ldo) {
    const int64_t total = m * k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / k;
        const int c = (int)(i - r * k);
        out[r * ldo + c] = src[(rows ? rows[r] : r) * ld + c] - __ldg(center + c);
    }
}

__global__ void rank_prepare_kernel(const float* __restrict__ emb, int64_t ld, const int64_t* __restrict__ tail_rows,
                                    const float* __restrict__ head_emb, int64_t ld_h,
                                    const int64_t* __restrict__ head_rows, const int64_t* __restrict__ target_pos,
                                    int n_heads, int dim, const float* __restrict__ tail_max_norm,
                                    const float* __restrict__ rec, const float* __restrict__ center,
                                    float* __restrict__ tau, float* __restrict__ thr) {
    const int lane = threadIdx.x & 31;
    const int head = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (head >= n_heads) return;
    const float* hrow = head_emb + (head_rows ? head_rows[head] : head) * ld_h;
    const int64_t tp = target_pos[head];
    const float* trow = emb + (tail_rows ? tail_rows[tp] : tp) * ld;
    double acc = 0.0, nh = 0.0, hc = 0.0;
    for (int c = lane; c < dim; c += 32) {
        const float a = __ldg(hrow + c), b = __ldg(trow + c);
        acc = fma((double)a, (double)b, acc);
        nh = fma((double)a, (double)a, nh);
        if (center) hc = fma((double)a, (double)__ldg(center + c), hc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(kFull, acc, o);
        nh += __shfl_xor_sync(kFull, nh, o);
        hc += __shfl_xor_sync(kFull, hc, o);
    }
    if (lane == 0) {
        const float t = (float)acc;                          // exact target score, rounded once (what finalize compares)
        const float ts = (float)(acc - hc);                  // the same in the GEMM's (shifted) frame
        // bound of the GEMM's error (norms of the tails: scaled units) + the rounding of the two fp32 numbers above
        const float e = kRankErr * (float)sqrt(nh) * (*tail_max_norm) * rec[2] * 1.0001f +
                        (fabsf(t) + fabsf(ts) + (float)fabs(hc)) * 2e-7f + 1e-30f;
        tau[head] = t;
        thr[2 * head] = ts - e;
        thr[2 * head + 1] = ts + e;
    }
}

__global__ void __launch_bounds__(256) rank_finalize_kernel(const float* __restrict__ emb, int64_t ld,
                                                            const int64_t* __restrict__ tail_rows,
                                                            const float* __restrict__ head_emb, int64_t ld_h,
                                                            const int64_t* __restrict__ head_rows,
                                                            const int64_t* __restrict__ target_pos,
                                                            const float* __restrict__ tau, const int* __restrict__ above,
                                                            const int* __restrict__ band_cnt, const int* __restrict__ band,
                                                            int cap, int n_tails, int dim, int64_t* __restrict__ ranks) {
    __shared__ __align__(16) float s_head[kMaxChunks * kBK];
    __shared__ int s_better[8];
    const int head = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int grp = lane >> 3, q = lane & 7;
    const float* hrow = head_emb + (head_rows ? head_rows[head] : head) * ld_h;
    for (int i = threadIdx.x; i < kMaxChunks * kBK; i += blockDim.x) s_head[i] = i < dim ? __ldg(hrow + i) : 0.f;
    __syncthreads();
    const float t = tau[head];
    const int64_t tp = target_pos[head];
    const int n_band = band_cnt[head];
    const bool overflow = n_band > cap;            // the band held more columns than the list: exact scan of every tail
    const int n = overflow ? n_tails : n_band;
    int better = 0;
    for (int c0 = 4 * warp; c0 < n; c0 += 4 * nwarps) {
        const int c = c0 + grp;
        const bool live = c < n;
        const int col = live ? (overflow ? c : band[(int64_t)head * cap + c]) : 0;
        const float* trow = emb + (tail_rows ? tail_rows[col] : col) * ld;
        const float s = exact_dot8(s_head, trow, dim, q, live);
        if (live && q == 0 && (s > t || (s == t && col < tp))) ++better;
    }
    better = (int)warp_sum((float)better);          // < 2^24 per warp
    if (lane == 0) s_better[warp] = better;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t tot = overflow ? 0 : above[head];
        for (int w = 0; w < nwarps; ++w) tot += s_better[w];
        ranks[head] = tot;
    }
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_score_index(const float* emb, int64_t ld, const int64_t* rows, int64_t m, int32_t dim,
                               const float* rec, uint16_t* hi, int64_t ld_hi, float* norms, float* max_norm,
                               void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(emb && rec && hi && norms && max_norm && m >= 0 && dim > 0, "null / bad argument");
    LKG_REQUIRE(ld_hi % 8 == 0 && ld_hi >= dim && aligned16(hi) && aligned16(emb), "hi plane rows must be 16-byte aligned");
    LKG_CUDA(cudaMemsetAsync(max_norm, 0, sizeof(float), stream));
    if (m == 0) return LKG_OK;
    int64_t blocks = (m * 32 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    score_index_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, stream>>>(emb, ld, rows, m, dim, rec, (__half*)hi,
                                                                              ld_hi, norms, max_norm);
    LKG_LAUNCH_CHECK("score_index_kernel");
    return LKG_OK;
}

extern "C" int lkg_shift_rows(const float* src, int64_t ld, const int64_t* rows, int64_t m, int32_t k, const float* center,
                              float* out, int64_t ld_out, void* stream_) {
    LKG_REQUIRE(src && center && out && m >= 0 && k > 0 && ld_out >= k, "bad shift arguments");
    if (m == 0) return LKG_OK;
    int64_t blocks = (m * k + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    shift_rows_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream_>>>(src, ld, rows, m, k, center,
                                                                                            out, ld_out);
    LKG_LAUNCH_CHECK("shift_rows_kernel");
    return LKG_OK;
}

extern "C" int lkg_rank_prepare(const float* emb, int64_t ld_emb, const int64_t* tail_rows, const float* head_emb,
                                int64_t ld_head_emb, const int64_t* head_rows, const int64_t* target_pos, int64_t n_heads,
                                int32_t dim, const float* tail_max_norm, const float* rec, const float* center,
                                float* tau, float* thr, void* stream_) {
    LKG_REQUIRE(emb && head_emb && target_pos && tail_max_norm && rec && tau && thr && n_heads >= 0 && dim > 0,
                "bad rank arguments");
    if (n_heads == 0) return LKG_OK;
    const int64_t blocks = (n_heads * 32 + 255) / 256;
    rank_prepare_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(emb, ld_emb, tail_rows, head_emb, ld_head_emb,
                                                                             head_rows, target_pos, (int)n_heads, dim,
                                                                             tail_max_norm, rec, center, tau, thr);
    LKG_LAUNCH_CHECK("rank_prepare_kernel");
    return LKG_OK;
}

extern "C" int lkg_rank_finalize(const float* emb, int64_t ld_emb, const int64_t* tail_rows, const float* head_emb,
                                 int64_t ld_head_emb, const int64_t* head_rows, const int64_t* target_pos,
                                 const float* tau, const int32_t* above, const int32_t* band_cnt, const int32_t* band,
                                 int32_t band_cap, int64_t n_heads, int64_t n_tails, int32_t dim, int64_t* ranks,
                                 void* stream_) {
    LKG_REQUIRE(emb && head_emb && target_pos && tau && above && band_cnt && band && ranks && band_cap > 0,
                "bad rank arguments");
    LKG_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= kMaxChunks * kBK && ld_emb % 4 == 0 && ld_head_emb % 4 == 0 &&
                    aligned16(emb) && aligned16(head_emb) && n_tails < (1ll << 31),
                "the rank path supports dim %% 4 == 0, dim <= %d, 16-byte aligned rows", kMaxChunks * kBK);
    if (n_heads == 0) return LKG_OK;
    rank_finalize_kernel<<<(unsigned)n_heads, 256, 0, (cudaStream_t)stream_>>>(emb, ld_emb, tail_rows, head_emb, ld_head_emb,
                                                                               head_rows, target_pos, tau, above, band_cnt,
                                                                               band, band_cap, (int)n_tails, dim, ranks);
    LKG_LAUNCH_CHECK("rank_finalize_kernel");
    return LKG_OK;
}

extern "C" int lkg_score_topk_workspace_bytes(int64_t n_heads, int32_t cap, int32_t sample_tiles, size_t* bytes) {
    LKG_REQUIRE(bytes && n_heads >= 0 && cap >= 2 && (cap & (cap - 1)) == 0, "cap must be a power of two");
    LKG_REQUIRE(sample_tiles >= 0 && sample_tiles <= 4096, "sample_tiles must be in [0, 4096]");
    // counters [n_heads, <= 160 streams], thresholds, candidate sublists (n_streams * cap_s <= cap), tile maxima
    *bytes = align_up((size_t)n_heads * 160 * 4) + align_up((size_t)n_heads * 4) + align_up((size_t)n_heads * cap * 4) +
             align_up((size_t)n_heads * sample_tiles * 4);
    return LKG_OK;
}

extern "C" int lkg_score_topk(const uint16_t* heads_hi, int64_t ld_heads_hi, const float* head_norms, int64_t n_heads,
                              const uint16_t* tails_hi, int64_t ld_tails_hi, const float* tail_max_norm,
                              int64_t n_tails, int32_t dim, const float* rec, const float* theta,
                              int64_t theta_stride, int32_t sample_tiles, const float* emb, int64_t ld_emb,
                              const float* head_emb, int64_t ld_head_emb, const int64_t* head_rows,
                              const int64_t* tail_rows, int32_t k, int32_t cap,
                              float* top_values, int64_t* top_cols, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(n_heads >= 0 && n_tails >= 1 && n_tails < (1ll << 31) - kBN, "bad shape");
    LKG_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= kMaxChunks * kBK,
                "the fused scoring path supports dim %% 4 == 0, dim <= %d (got %d)", kMaxChunks * kBK, dim);
    LKG_REQUIRE(ld_emb % 4 == 0 && aligned16(emb) && (!head_emb || (ld_head_emb % 4 == 0 && aligned16(head_emb))),
                "embedding rows must be 16-byte aligned");
    LKG_REQUIRE(k >= 1 && cap >= 2 * k && (cap & (cap - 1)) == 0 && cap <= 16384,
                "cap must be a power of two in [2k, 16384]");
    if (n_heads == 0) return LKG_OK;
    LKG_REQUIRE(heads_hi && head_norms && tails_hi && tail_max_norm && rec && emb && top_values && top_cols && workspace,
                "null argument");
    char* ws = static_cast<char*>(workspace);
    int* cnt = reinterpret_cast<int*>(ws);
    float* thr = reinterpret_cast<float*>(ws + align_up((size_t)n_heads * 160 * 4));
    int* cand = reinterpret_cast<int*>(ws + align_up((size_t)n_heads * 160 * 4) + align_up((size_t)n_heads * 4));
    float* tilemax = reinterpret_cast<float*>(ws + align_up((size_t)n_heads * 160 * 4) + align_up((size_t)n_heads * 4) +
                                              align_up((size_t)n_heads * cap * 4));
    if (theta || sample_tiles <= 0) {
        score_threshold_kernel<<<(int)((n_heads + 255) / 256), 256, 0, stream>>>(theta, theta_stride, (int)n_heads,
                                                                                head_norms, tail_max_norm, rec, dim, thr);
        LKG_LAUNCH_CHECK("score_threshold_kernel");
    }
    const int n_tiles_all = (int)((n_tails + kBN - 1) / kBN);
    int n_st = 0, st_stride = 1;
    if (!theta && sample_tiles > 0) {
        LKG_REQUIRE(sample_tiles <= 4096, "sample_tiles must be <= 4096");
        st_stride = n_tiles_all / sample_tiles > 0 ? n_tiles_all / sample_tiles : 1;
        n_st = (n_tiles_all + st_stride - 1) / st_stride;
        if (n_st > sample_tiles) n_st = sample_tiles;
    }

    const int sms = sm_count() < 160 ? sm_count() : 160;
    const int max_pairs = sms;                      // heads per launch <= 256 * SMs
    auto kern = score_filter_kernel;
    LKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFilterSmem));
    const size_t fsmem = (size_t)cap * (sizeof(unsigned long long) + sizeof(int));
    LKG_CUDA(cudaFuncSetAttribute(score_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
    for (int64_t h0 = 0; h0 < n_heads; h0 += (int64_t)max_pairs * 2 * kBM) {
        const int nh = (int)((n_heads - h0) < (int64_t)max_pairs * 2 * kBM ? (n_heads - h0) : (int64_t)max_pairs * 2 * kBM);
        FilterParams p{};
        if (int rc = make_map_2d(&p.a_map, heads_hi + h0 * ld_heads_hi, nh, dim, ld_heads_hi)) return rc;
        if (int rc = make_map_2d(&p.b_map, tails_hi, n_tails, dim, ld_tails_hi)) return rc;
        p.n_heads = nh;
        p.n_tails = (int)n_tails;
        p.n_chunks = (dim + kBK - 1) / kBK;
        p.n_pairs = (nh + 2 * kBM - 1) / (2 * kBM);
        p.thr = thr + h0;
        // pass 0 (optional): tile maxima over the sampled tiles -> thresholds; pass 1: filter over all tiles
        int streams = 1;
        for (int pass = (n_st > 0 ? 0 : 1); pass < 2; ++pass) {
            p.n_tiles = pass == 0 ? n_st : n_tiles_all;
            p.tile_stride = pass == 0 ? st_stride : 1;
            p.tilemax = pass == 0 ? tilemax + h0 * n_st : nullptr;
            streams = sms / p.n_pairs;
            if (streams > p.n_tiles) streams = p.n_tiles;
            p.cnt = cnt + h0 * 160;
            p.cand = cand + h0 * cap;
            p.cap_s = cap / streams;
            kern<<<p.n_pairs * streams, kThreads, kFilterSmem, stream>>>(p);
            LKG_LAUNCH_CHECK("score_filter_kernel");
            if (pass == 0) {
                int p2 = 1;
                while (p2 < n_st) p2 <<= 1;
                score_sample_threshold_kernel<<<(unsigned)nh, 256, p2 * sizeof(float), stream>>>(
                    p.tilemax, n_st, k, head_norms + h0, tail_max_norm, dim, thr + h0);
                LKG_LAUNCH_CHECK("score_sample_threshold_kernel");
            }
        }

        FinalParams f{};
        f.emb = emb;
        f.ld_emb = ld_emb;
        f.head_emb = head_emb ? head_emb : emb;
        f.ld_head_emb = head_emb ? ld_head_emb : ld_emb;
        f.head_rows = head_rows ? head_rows + h0 : nullptr;
        f.head_row_base = head_rows ? 0 : h0;
        f.tail_rows = tail_rows;
        f.n_heads = nh;
        f.n_tails = (int)n_tails;
        f.dim = dim;
        f.k = k;
        f.cap = cap;
        f.n_streams = streams;
        f.cap_s = cap / streams;
        f.cnt = cnt + h0 * 160;
        f.cand = cand + h0 * cap;
        f.top_val = top_values + h0 * k;
        f.top_col = top_cols + h0 * k;
        score_finalize_kernel<<<(unsigned)nh, kFinalThreads, fsmem, stream>>>(f);
        LKG_LAUNCH_CHECK("score_finalize_kernel");
    }
    return LKG_OK;
}
