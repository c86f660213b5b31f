// Attention update: per-triple logit, duplicate-(h,t) merge, row softmax -> A_in values.
//
// Replaces LiteralKG.update_attention / update_attention_batch (model.py:430-471):
//     v(h,r,t) = sum_d e_t[d] * tanh(e_h[d] + e_r[d])        on the RAW entity / relation tables
//     A[h,t]   = softmax_t( sum_{r : (h,r,t) in KG} v(h,r,t) )
// The reference loops over relations in Python (where / gather / tanh / sum), concatenates an
// un-coalesced COO tensor, copies it to the host, runs torch.sparse.softmax there (coalesce = sort +
// duplicate sum) and copies the result back.  Here one warp owns one head row of the plan:
//   * tanh(a + b) = 1 - 2 / (1 + exp(2a) exp(2b)): exp(2 e_h) is computed once per row (registers) and
//     exp(2 e_r) once per launch (R x D table in the workspace, L1/L2 resident), so a (head, relation)
//     weight vector costs one multiply, one add, one MUFU.RCP and one FMA per element instead of a
//     ~25-instruction tanh -- with R = 64 uniform relations almost every triple starts a new run and
//     the old per-run tanh made the kernel issue bound (ncu: 52 % issue slots at 22 % occupancy);
//   * triples are visited in (h, r, t) order, so the weight vector is still shared by a whole run;
//   * each tail row e_t is fetched exactly once with 128-bit streaming loads, kUnroll tails in flight;
//   * the logit is accumulated into its (h,t) pair slot (att_seg), which sums duplicates -- in shared
//     memory for rows of up to kSegCap pairs, in the output array itself for longer rows;
//   * the softmax over the row's pair slots runs in the same warp;
//   * rows are taken in the plan's degree-descending order (row_order): the 4096-neighbour rows of a
//     power-law graph start first instead of forming the kernel's tail.
// Envelope: exact formula for |e| <= 40 (beyond that exp(2e) is clamped; tanh is saturated anyway).
// HBM bound: algorithmic bytes = E*(4 tail + 4 rel + 4 seg + 4D) + N*(4D + 8 + 4 order) + nnz*4.
#include "common.cuh"

namespace lkg {
namespace {

constexpr int kUnroll = 4;
constexpr int kWarps = 8;
constexpr int kSegCap = 128;                     // pair slots per warp kept in shared memory
constexpr float kTwoLog2e = 2.8853900817779268f; // exp(2x) = exp2(x * 2 log2 e)

__device__ __forceinline__ float exp2x(float x) { return exp2f(kTwoLog2e * fminf(fmaxf(x, -40.f), 40.f)); }

__global__ void rel_exp_kernel(const float* __restrict__ rel, int64_t ld_rel, int n_rel, int dim,
                               float* __restrict__ out) {
    const int total = n_rel * dim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / dim, d = i - r * dim;
        out[i] = exp2x(rel[(int64_t)r * ld_rel + d]);
    }
}

// 1 - 2 / (1 + p): tanh of the summed arguments; p = inf -> 1, p = 0 -> -1
__device__ __forceinline__ float tanh_from_exp(float p) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + p));
    return fmaf(-2.f, r, 1.f);
}

template <int S>  // S = float4 slots per lane: dim <= 128 * S
__global__ void __launch_bounds__(kWarps * 32)
attn_update_kernel(lkg_graph g, const float* __restrict__ ent, int64_t ld_ent,
                   const float* __restrict__ rel_exp /* [R, 4 * nvec] */, int nvec /* dim/4 */,
                   float* __restrict__ val, int* __restrict__ row_counter) {
    __shared__ float s_logit[kWarps][kSegCap];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int row0 = (int)g.row_begin, n_rows = (int)(g.row_end - g.row_begin);

    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(row_counter, 1);
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= n_rows) break;
        const int row = g.row_order ? __ldg(g.row_order + idx) : row0 + idx;
        const int e0 = g.att_rowptr[row], e1 = g.att_rowptr[row + 1];
        if (e0 == e1) continue;
        const int u0 = g.rowptr[row], nu = g.rowptr[row + 1] - u0;
        float* logit = nu <= kSegCap ? s_logit[warp] : val + u0;      // slot of pair u: logit[u - u0]
        for (int i = lane; i < nu; i += 32) logit[i] = 0.f;

        float4 eh[S], w[S];
        const float* hrow = ent + (int64_t)row * ld_ent;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int v = lane + 32 * s;
            eh[s] = v < nvec ? __ldg(reinterpret_cast<const float4*>(hrow) + v) : make_float4(0, 0, 0, 0);
            eh[s].x = exp2x(eh[s].x);
            eh[s].y = exp2x(eh[s].y);
            eh[s].z = exp2x(eh[s].z);
            eh[s].w = exp2x(eh[s].w);
            w[s] = make_float4(0, 0, 0, 0);
        }
        __syncwarp();   // zeroing of the slots visible before lane 0 accumulates

        int cur_rel = -1;
        for (int e = e0; e < e1; e += kUnroll) {
            int tl[kUnroll], rl[kUnroll], sg[kUnroll];
            float4 et[kUnroll][S];
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const bool live = e + j < e1;
                tl[j] = live ? __ldg(g.att_tail + e + j) : 0;
                rl[j] = live ? __ldg(g.att_rel + e + j) : -1;
                sg[j] = live ? __ldg(g.att_seg + e + j) : 0;
            }
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const float* trow = ent + (int64_t)tl[j] * ld_ent;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int v = lane + 32 * s;
                    et[j][s] = (rl[j] >= 0 && v < nvec) ? ldg_stream4(trow + 4 * v) : make_float4(0, 0, 0, 0);
                }
            }
            float part[kUnroll];
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                if (rl[j] >= 0 && rl[j] != cur_rel) {   // warp-uniform
                    cur_rel = rl[j];
                    const float4* rrow = reinterpret_cast<const float4*>(rel_exp) + (int64_t)cur_rel * nvec;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const int v = lane + 32 * s;
                        if (v < nvec) {
                            const float4 er = __ldg(rrow + v);
                            w[s].x = tanh_from_exp(eh[s].x * er.x);
                            w[s].y = tanh_from_exp(eh[s].y * er.y);
                            w[s].z = tanh_from_exp(eh[s].z * er.z);
                            w[s].w = tanh_from_exp(eh[s].w * er.w);
                        }
                    }
                }
                float p = 0.f;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    p = fmaf(et[j][s].x, w[s].x, p);
                    p = fmaf(et[j][s].y, w[s].y, p);
                    p = fmaf(et[j][s].z, w[s].z, p);
                    p = fmaf(et[j][s].w, w[s].w, p);
                }
                part[j] = p;
            }
            // four warp sums in 6 shuffles: halve the number of live values on the first two steps
            static_assert(kUnroll == 4, "the folded reduction below is written for 4 values");
            const bool up16 = lane & 16, up8 = lane & 8;
            float k0 = (up16 ? part[2] : part[0]) + __shfl_xor_sync(kFull, up16 ? part[0] : part[2], 16);
            float k1 = (up16 ? part[3] : part[1]) + __shfl_xor_sync(kFull, up16 ? part[1] : part[3], 16);
            float k = (up8 ? k1 : k0) + __shfl_xor_sync(kFull, up8 ? k0 : k1, 8);
            k += __shfl_xor_sync(kFull, k, 4);
            k += __shfl_xor_sync(kFull, k, 2);
            k += __shfl_xor_sync(kFull, k, 1);     // lanes with (lane >> 3) == j hold the logit of triple j
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const float v = __shfl_sync(kFull, k, 8 * j);
                if (lane == 0 && rl[j] >= 0) logit[sg[j] - u0] += v;   // sequential: duplicates of a pair add up in order
            }
        }
        __syncwarp();

        // softmax over the row's unique pairs
        float m = -INFINITY;
        for (int i = lane; i < nu; i += 32) m = fmaxf(m, logit[i]);
        m = warp_max(m);
        float sum = 0.f;
        for (int i = lane; i < nu; i += 32) {
            const float ex = __expf(logit[i] - m);
            logit[i] = ex;
            sum += ex;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int i = lane; i < nu; i += 32) val[u0 + i] = logit[i] * inv;
        __syncwarp();   // the shared slots are reused by the next row
    }
}

constexpr size_t kCounterBytes = 256;

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_attn_workspace_bytes(int32_t n_relations, int32_t dim, size_t* bytes) {
    LKG_REQUIRE(bytes != nullptr && n_relations > 0 && dim > 0, "bad workspace query");
    *bytes = kCounterBytes + (size_t)n_relations * dim * sizeof(float);
    return LKG_OK;
}

extern "C" int lkg_attn_update(const lkg_graph* g, const float* entity, int64_t ld_entity,
                               const float* relation, int64_t ld_relation, int32_t dim,
                               float* values, void* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(g && entity && relation && workspace, "null argument");
    LKG_REQUIRE(g->nnz == 0 || values != nullptr, "values is null");
    LKG_REQUIRE(g->row_begin >= 0 && g->row_begin <= g->row_end && g->row_end <= g->n_entities, "bad row range");
    LKG_REQUIRE(dim > 0 && dim % 4 == 0, "dim must be a positive multiple of 4 (got %d)", dim);
    LKG_REQUIRE(ld_entity % 4 == 0 && aligned16(entity) && aligned16(workspace),
                "entity rows and the workspace must be 16-byte aligned");
    if (dim > 512) LKG_FAIL(LKG_ERR_UNSUPPORTED, "attention dim %d > 512", dim);
    if (g->n_edges == 0) return LKG_OK;
    int* counter = static_cast<int*>(workspace);
    float* rel_exp = reinterpret_cast<float*>(static_cast<char*>(workspace) + kCounterBytes);
    LKG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    const int total = g->n_relations * dim;
    rel_exp_kernel<<<(total + 255) / 256, 256, 0, stream>>>(relation, ld_relation, g->n_relations, dim, rel_exp);
    LKG_LAUNCH_CHECK("rel_exp_kernel");
    const int nvec = dim / 4;
    const int slots = (nvec + 31) / 32;
    const int block = kWarps * 32;
    const int grid = sm_count() * 8;
    switch (slots) {
        case 1: attn_update_kernel<1><<<grid, block, 0, stream>>>(*g, entity, ld_entity, rel_exp, nvec, values, counter); break;
        case 2: attn_update_kernel<2><<<grid, block, 0, stream>>>(*g, entity, ld_entity, rel_exp, nvec, values, counter); break;
        case 3: attn_update_kernel<3><<<grid, block, 0, stream>>>(*g, entity, ld_entity, rel_exp, nvec, values, counter); break;
        default: attn_update_kernel<4><<<grid, block, 0, stream>>>(*g, entity, ld_entity, rel_exp, nvec, values, counter); break;
    }
    LKG_LAUNCH_CHECK("attn_update_kernel");
    return LKG_OK;
}
