// Attention update: per-triple logit, duplicate-(h,t) merge, row softmax -> A_in values.
//
// Replaces LiteralKG.update_attention / update_attention_batch (model.py:430-471):
//     v(h,r,t) = sum_d e_t[d] * tanh(e_h[d] + e_r[d])        on the RAW entity / relation tables
//     A[h,t]   = softmax_t( sum_{r : (h,r,t) in KG} v(h,r,t) )
// The reference loops over relations in Python (where / gather / tanh / sum), concatenates an
// un-coalesced COO tensor, copies it to the host, runs torch.sparse.softmax there (coalesce = sort +
// duplicate sum) and copies the result back.  Here one warp owns one head row of the plan:
//   * triples are visited in (h, r, t) order, so w = tanh(e_h + e_r) is computed once per
//     (head, relation) run and kept in registers;
//   * each tail row e_t is fetched exactly once with 128-bit streaming loads, UNROLL tails in flight;
//   * the logit is accumulated into its (h,t) pair slot (att_seg), which sums duplicates;
//   * the softmax over the row's pair slots runs in the same warp.
// HBM bound: algorithmic bytes = E*(4 tail + 4 rel + 4 seg + 4D) + N*(4D + 8) + nnz*4.
#include "common.cuh"

namespace lkg {
namespace {

constexpr int kUnroll = 4;

template <int S>  // S = float4 slots per lane: dim <= 128 * S
__global__ void __launch_bounds__(256)
attn_update_kernel(lkg_graph g, const float* __restrict__ ent, int64_t ld_ent,
                   const float* __restrict__ rel, int64_t ld_rel, int nvec /* dim/4 */,
                   float* __restrict__ val, int* __restrict__ row_counter) {
    const int lane = threadIdx.x & 31;
    const int row0 = (int)g.row_begin, n = (int)g.row_end;

    for (;;) {
        int row = 0;
        if (lane == 0) row = row0 + atomicAdd(row_counter, 1);
        row = __shfl_sync(kFull, row, 0);
        if (row >= n) break;
        const int e0 = g.att_rowptr[row], e1 = g.att_rowptr[row + 1];
        if (e0 == e1) continue;
        const int u0 = g.rowptr[row], u1 = g.rowptr[row + 1];
        for (int u = u0 + lane; u < u1; u += 32) val[u] = 0.f;

        float4 eh[S], w[S];
        const float* hrow = ent + (int64_t)row * ld_ent;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int v = lane + 32 * s;
            eh[s] = v < nvec ? __ldg(reinterpret_cast<const float4*>(hrow) + v) : make_float4(0, 0, 0, 0);
            w[s] = make_float4(0, 0, 0, 0);
        }
        __syncwarp();   // zeroing of val[] visible before lane 0 accumulates

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
                    const float* rrow = rel + (int64_t)cur_rel * ld_rel;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const int v = lane + 32 * s;
                        if (v < nvec) {
                            const float4 er = __ldg(reinterpret_cast<const float4*>(rrow) + v);
                            w[s].x = tanh_acc(eh[s].x + er.x);
                            w[s].y = tanh_acc(eh[s].y + er.y);
                            w[s].z = tanh_acc(eh[s].z + er.z);
                            w[s].w = tanh_acc(eh[s].w + er.w);
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
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < kUnroll; ++j) part[j] += __shfl_xor_sync(kFull, part[j], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < kUnroll; ++j)
                    if (rl[j] >= 0) val[sg[j]] += part[j];   // sequential: duplicates of a pair add up
            }
        }
        __syncwarp();

        // softmax over the row's unique pairs
        float m = -INFINITY;
        for (int u = u0 + lane; u < u1; u += 32) m = fmaxf(m, val[u]);
        m = warp_max(m);
        float sum = 0.f;
        for (int u = u0 + lane; u < u1; u += 32) {
            const float ex = __expf(val[u] - m);
            val[u] = ex;
            sum += ex;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int u = u0 + lane; u < u1; u += 32) val[u] *= inv;
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_attn_workspace_bytes(size_t* bytes) {
    LKG_REQUIRE(bytes != nullptr, "bytes is null");
    *bytes = 256;
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
    LKG_REQUIRE(ld_entity % 4 == 0 && ld_relation % 4 == 0 && aligned16(entity) && aligned16(relation),
                "entity / relation rows must be 16-byte aligned");
    if (dim > 512) LKG_FAIL(LKG_ERR_UNSUPPORTED, "attention dim %d > 512", dim);
    if (g->n_edges == 0) return LKG_OK;
    int* counter = static_cast<int*>(workspace);
    LKG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    const int nvec = dim / 4;
    const int slots = (nvec + 31) / 32;
    const int block = 256;
    const int grid = sm_count() * 6;
    switch (slots) {
        case 1: attn_update_kernel<1><<<grid, block, 0, stream>>>(*g, entity, ld_entity, relation, ld_relation, nvec, values, counter); break;
        case 2: attn_update_kernel<2><<<grid, block, 0, stream>>>(*g, entity, ld_entity, relation, ld_relation, nvec, values, counter); break;
        case 3: attn_update_kernel<3><<<grid, block, 0, stream>>>(*g, entity, ld_entity, relation, ld_relation, nvec, values, counter); break;
        default: attn_update_kernel<4><<<grid, block, 0, stream>>>(*g, entity, ld_entity, relation, ld_relation, nvec, values, counter); break;
    }
    LKG_LAUNCH_CHECK("attn_update_kernel");
    return LKG_OK;
}
