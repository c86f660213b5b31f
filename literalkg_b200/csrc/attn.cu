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
//   * each tail row e_t is fetched exactly once: a row's (tail, relation, pair) indices are loaded 32 at a time in one
//     coalesced step and the 1 200-byte tail rows stream through a per-warp shared-memory ring filled by
//     cp.async.bulk (one TMA-unit copy per row, completion on an mbarrier), kRing rows in flight per warp without
//     holding registers -- the register-staged version kept 4 rows in flight behind a dependent index -> gather
//     chain and sat at 71 % of the HBM peak with long_scoreboard as the top stall;
//   * the logit is accumulated into its (h,t) pair slot (att_seg), which sums duplicates -- in shared
//     memory for rows of up to kSegCap pairs, in the output array itself for longer rows;
//   * the softmax over the row's pair slots runs in the same warp;
//   * rows are taken in the plan's degree-descending order (row_order): the 4096-neighbour rows of a
//     power-law graph start first instead of forming the kernel's tail.
// Envelope: exact formula for |e| <= 40 (beyond that exp(2e) is clamped; tanh is saturated anyway).
// HBM bound: algorithmic bytes = E*(4 tail + 4 rel + 4 seg + 4D) + N*(4D + 8 + 4 order) + nnz*4.
#include <stdlib.h>

#include "common.cuh"

namespace lkg {
namespace {

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

// S = float4 slots per lane (dim <= 128 * S); kRing = tail rows in flight per warp (bulk copies into shared memory)
//
// Software pipeline across rows (every load below is a global round trip of 1-2 us under load, and a warp spends
// only ~5 us streaming an average row, so a chain of them per row would dominate):
//   during row i     the counter fetch for row i+2, then the schedule record of row i+2;
//                    the first index chunk and the head row e_h of row i+1 (its record arrived during row i-1)
//   at row i start   everything is in registers: the ring copies of the first chunk are issued immediately
template <int S, int kRing, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
attn_update_kernel(lkg_graph g, const float* __restrict__ ent, int64_t ld_ent,
                   const float* __restrict__ rel_exp /* [R, 4 * nvec] */, int nvec /* dim/4 */,
                   float* __restrict__ val, int* __restrict__ row_counter) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t row_bytes = (uint32_t)nvec * 16u;
    // per warp: kRing row slots, kRing mbarriers, kSegCap logit slots
    uint8_t* ring = smem_raw + (size_t)warp * kRing * row_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWarps * kRing * row_bytes) + warp * kRing;
    float* s_logit = reinterpret_cast<float*>(smem_raw + (size_t)kWarps * kRing * row_bytes + kWarps * kRing * 8) +
                     warp * kSegCap;
    if (lane < kRing) ring_bar_init(&bars[lane]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int n_rows = g.n_sched > 0 ? (int)g.n_sched : (int)(g.row_end - g.row_begin);   // schedule records
    uint32_t issued = 0, consumed = 0;              // running counts over the kernel: slot = count % kRing,
                                                    // parity of a slot's use = (count / kRing) & 1
    const int4* sched = reinterpret_cast<const int4*>(g.row_sched);

    struct Rec { int row, e0, e1, u0, u1, nseg, ticket; };
    auto fetch_idx = [&]() {                         // dynamic balance: the heaviest rows are first in the schedule
        int i = 0;
        if (lane == 0) i = atomicAdd(row_counter, 1);
        return i;                                    // valid in lane 0 until broadcast
    };
    auto load_rec = [&](int i) {
        Rec r{0, 0, 0, 0, 0, 0, 0};
        if (i < n_rows) {
            const int4 a = __ldg(sched + 2 * i);
            const int4 b = __ldg(sched + 2 * i + 1);
            r.row = a.x; r.e0 = a.y; r.e1 = a.z; r.u0 = a.w;
            r.u1 = b.x; r.nseg = b.y; r.ticket = b.z;
            if (r.nseg > 0) {      // a piece of a segmented row: the logit slots / softmax span the whole row
                r.u0 = __ldg(g.rowptr + r.row);
                r.u1 = __ldg(g.rowptr + r.row + 1);
            }
        }
        return r;
    };
    // first chunk of a row: this lane's (tail, relation, pair) and the raw head row
    struct Head { int tail, rel, seg; float4 eh[S]; };
    auto load_head = [&](const Rec& r) {
        Head h;
        const int cn = min(32, r.e1 - r.e0);
        h.tail = lane < cn ? __ldg(g.att_tail + r.e0 + lane) : 0;
        h.rel = lane < cn ? __ldg(g.att_rel + r.e0 + lane) : -1;
        h.seg = lane < cn ? __ldg(g.att_seg + r.e0 + lane) : 0;
        const float* hrow = ent + (int64_t)r.row * ld_ent;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int v = lane + 32 * s;
            h.eh[s] = (v < nvec && r.e1 > r.e0) ? __ldg(reinterpret_cast<const float4*>(hrow) + v) : make_float4(0, 0, 0, 0);
        }
        return h;
    };

    // pipeline fill: rows 0 and 1 of this warp
    int idx = __shfl_sync(kFull, fetch_idx(), 0);
    Rec rec = load_rec(idx);
    int idx1 = __shfl_sync(kFull, fetch_idx(), 0);
    Rec rec1 = load_rec(idx1);
    Head head = load_head(rec);
    int pre_issued = 0;                              // copies of this row's first chunk issued during the previous row

    while (idx < n_rows && rec.e1 > rec.e0) {        // the schedule is sorted by triple count: an empty row ends it
        int idx2 = fetch_idx();                      // row i+2: broadcast and record load after the ring is filled
        const int row = rec.row, e0 = rec.e0, e1 = rec.e1, u0 = rec.u0, nu = rec.u1 - rec.u0;
        (void)row;
        const bool segmented = rec.nseg > 0;          // the row's pieces share the slots in `val` (zeroed by the host)
        const bool in_smem = !segmented && nu <= kSegCap;
        float* logit = in_smem ? s_logit : val + u0;                   // slot of pair u: logit[u - u0]
        if (!segmented)
            for (int i = lane; i < nu; i += 32) logit[i] = 0.f;
        float4 eh[S], w[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            eh[s] = head.eh[s];
            w[s] = make_float4(0, 0, 0, 0);
        }
        __syncwarp();   // zeroing of the slots visible before the logits are written

        Rec rec2{0, 0, 0, 0, 0, 0, 0};
        Head head1;
        int cur_rel = -1;
        // the row's triples in chunks of 32: one coalesced load of (tail, relation, pair) per chunk, then the tail rows
        // stream through the ring, kRing bulk copies in flight
        for (int c0 = e0; c0 < e1; c0 += 32) {
            const int cn = min(32, e1 - c0);
            int my_tail, my_rel, my_seg;
            if (c0 == e0) {
                my_tail = head.tail; my_rel = head.rel; my_seg = head.seg;
            } else {
                my_tail = lane < cn ? __ldg(g.att_tail + c0 + lane) : 0;
                my_rel = lane < cn ? __ldg(g.att_rel + c0 + lane) : -1;
                my_seg = lane < cn ? __ldg(g.att_seg + c0 + lane) : 0;
            }
            // prologue: fill the ring
            const int first = min(cn, kRing);
            const int done = c0 == e0 ? pre_issued : 0;          // the first chunk's copies may already be in flight
            if (lane >= done && lane < first) {
                const uint32_t slot = (issued + lane - done) % kRing;
                ring_issue(ring + slot * row_bytes, ent + (int64_t)my_tail * ld_ent, row_bytes, &bars[slot]);
            }
            issued += first - done;
            if (c0 == e0) {
                // with this row's copies in flight: the loads of the rows to come, then exp(2 e_h)
                idx2 = __shfl_sync(kFull, idx2, 0);
                rec2 = load_rec(idx2);
                head1 = load_head(rec1);
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    eh[s].x = exp2x(eh[s].x);
                    eh[s].y = exp2x(eh[s].y);
                    eh[s].z = exp2x(eh[s].z);
                    eh[s].w = exp2x(eh[s].w);
                }
            }
            for (int j0 = 0; j0 < cn; j0 += 4) {
                float part[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = j0 + j;
                    const int rl = __shfl_sync(kFull, my_rel, e & 31);
                    float p = 0.f;
                    if (e < cn) {                               // warp-uniform
                        if (rl != cur_rel) {
                            cur_rel = rl;
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
                        const uint32_t slot = consumed % kRing;
                        ring_wait(&bars[slot], (consumed / kRing) & 1);
                        const float4* trow = reinterpret_cast<const float4*>(ring + slot * row_bytes);
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const int v = lane + 32 * s;
                            if (v < nvec) {
                                const float4 et = trow[v];
                                p = fmaf(et.x, w[s].x, p);
                                p = fmaf(et.y, w[s].y, p);
                                p = fmaf(et.z, w[s].z, p);
                                p = fmaf(et.w, w[s].w, p);
                            }
                        }
                        ++consumed;
                        __syncwarp();                            // every lane has read the slot before it is refilled
                        const int nxt = e + kRing;               // keep kRing copies in flight
                        if (nxt < cn) {
                            if (lane == (nxt & 31))
                                ring_issue(ring + slot * row_bytes, ent + (int64_t)my_tail * ld_ent, row_bytes, &bars[slot]);
                            ++issued;
                        }
                    }
                    part[j] = p;
                }
                // four warp sums in 6 shuffles: halve the number of live values on the first two steps
                const bool up16 = lane & 16, up8 = lane & 8;
                float k0 = (up16 ? part[2] : part[0]) + __shfl_xor_sync(kFull, up16 ? part[0] : part[2], 16);
                float k1 = (up16 ? part[3] : part[1]) + __shfl_xor_sync(kFull, up16 ? part[1] : part[3], 16);
                float k = (up8 ? k1 : k0) + __shfl_xor_sync(kFull, up8 ? k0 : k1, 8);
                k += __shfl_xor_sync(kFull, k, 4);
                k += __shfl_xor_sync(kFull, k, 2);
                k += __shfl_xor_sync(kFull, k, 1);     // lanes with (lane >> 3) == j hold the logit of triple j0 + j
                // lane 8 j owns triple j0 + j: a triple that is alone in its (h,t) pair stores its logit; the rare
                // triples of multi-relation pairs (bit 31 of att_seg) are added (the slots start at zero).  A lane-0
                // load + add + store per triple was a dependent chain that cost 19 % of the stall samples.
                const int mine = j0 + (lane >> 3);
                const int sg = __shfl_sync(kFull, my_seg, mine & 31);
                if ((lane & 7) == 0 && mine < cn) {
                    float* slot = logit + ((sg & 0x7fffffff) - u0);
                    if (sg >= 0) *slot = k;
                    else atomicAdd(slot, k);
                }
            }
        }
        // the ring is empty: start the next row's first copies now, they land during the softmax and the row switch
        pre_issued = 0;
        if (idx1 < n_rows && rec1.e1 > rec1.e0) {
            pre_issued = min(kRing, min(32, rec1.e1 - rec1.e0));
            if (lane < pre_issued) {
                const uint32_t slot = (issued + lane) % kRing;
                ring_issue(ring + slot * row_bytes, ent + (int64_t)head1.tail * ld_ent, row_bytes, &bars[slot]);
            }
            issued += pre_issued;
        }
        if (!in_smem) __threadfence();          // the long-row reductions are performed before other lanes read them
        __syncwarp();
        bool finish = true;
        if (segmented) {                        // the piece that arrives last runs the row's softmax
            int t = 0;
            if (lane == 0) {
                t = atomicAdd(g.seg_tickets + rec.ticket, 1);
                if (t == rec.nseg - 1) g.seg_tickets[rec.ticket] = 0;      // ready for the next launch
            }
            t = __shfl_sync(kFull, t, 0);
            finish = t == rec.nseg - 1;
            __threadfence();
        }

        // softmax over the row's unique pairs (__ldcg: the long-row logits were reduced in L2, bypass L1)
        if (finish) {
        float m = -INFINITY;
        for (int i = lane; i < nu; i += 32) m = fmaxf(m, in_smem ? logit[i] : __ldcg(logit + i));
        m = warp_max(m);
        float sum = 0.f;
        for (int i = lane; i < nu; i += 32) {
            const float ex = __expf((in_smem ? logit[i] : __ldcg(logit + i)) - m);
            logit[i] = ex;
            sum += ex;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int i = lane; i < nu; i += 32) val[u0 + i] = logit[i] * inv;
        }
        __syncwarp();   // the shared slots are reused by the next row
        idx = idx1; rec = rec1; head = head1;
        idx1 = idx2; rec1 = rec2;
    }
}

// ---- relation-projected attention: logits per (head, relation) run -------------------------------------------------
// One warp per run: its weight vector u (run_w row, coalesced) stays in registers, the run's tail rows are gathered with
// streaming 128-bit loads (4 rows in flight), one warp reduction per triple.
template <int S>
__global__ void __launch_bounds__(256) attn_run_logits_kernel(const int* __restrict__ run_ptr,
                                                              const int* __restrict__ run_slot, int64_t n_runs,
                                                              const int* __restrict__ att_tail,
                                                              const float* __restrict__ ent, int64_t ld_ent, int nvec,
                                                              const float* __restrict__ run_w, int64_t ld_w,
                                                              float* __restrict__ logits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t run = warp; run < n_runs; run += nwarps) {
        const int e0 = __ldg(run_ptr + run), e1 = __ldg(run_ptr + run + 1);
        const float4* wrow = reinterpret_cast<const float4*>(run_w + (int64_t)__ldg(run_slot + run) * ld_w);
        float4 w[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int v = lane + 32 * s;
            w[s] = v < nvec ? __ldg(wrow + v) : make_float4(0, 0, 0, 0);
        }
        for (int c0 = e0; c0 < e1; c0 += 4) {
            float4 t[4][S];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool live = c0 + j < e1;
                const float* trow = ent + (int64_t)(live ? __ldg(att_tail + c0 + j) : 0) * ld_ent;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int v = lane + 32 * s;
                    t[j][s] = (live && v < nvec) ? ldg_stream4(trow + 4 * v) : make_float4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float p = 0.f;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    p = fmaf(t[j][s].x, w[s].x, p);
                    p = fmaf(t[j][s].y, w[s].y, p);
                    p = fmaf(t[j][s].z, w[s].z, p);
                    p = fmaf(t[j][s].w, w[s].w, p);
                }
                p = warp_sum(p);
                if (lane == 0 && c0 + j < e1) logits[c0 + j] = p;
            }
        }
    }
}

// in-place softmax of every CSR row: one warp per row, three passes over the row's slots
__global__ void __launch_bounds__(256) row_softmax_kernel(const int* __restrict__ rowptr, int64_t n_rows,
                                                          float* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n_rows; row += nwarps) {
        const int u0 = __ldg(rowptr + row), u1 = __ldg(rowptr + row + 1);
        float m = -INFINITY;
        for (int i = u0 + lane; i < u1; i += 32) m = fmaxf(m, val[i]);
        m = warp_max(m);
        float sum = 0.f;
        for (int i = u0 + lane; i < u1; i += 32) {
            const float ex = __expf(val[i] - m);
            val[i] = ex;
            sum += ex;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int i = u0 + lane; i < u1; i += 32) val[i] *= inv;
    }
}

inline size_t attn_smem_bytes(int nvec, int ring) {
    return (size_t)kWarps * ring * nvec * 16 + kWarps * ring * 8 + kWarps * kSegCap * 4;
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
    LKG_REQUIRE(g->row_sched != nullptr && aligned16(g->row_sched), "the plan has no row schedule (lkg_plan_build row_sched)");
    if (dim > 512) LKG_FAIL(LKG_ERR_UNSUPPORTED, "attention dim %d > 512", dim);
    if (g->n_edges == 0) return LKG_OK;
    int* counter = static_cast<int*>(workspace);
    float* rel_exp = reinterpret_cast<float*>(static_cast<char*>(workspace) + kCounterBytes);
    LKG_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), stream));
    if (g->n_sched > 0 && g->seg_tickets) {
        // the pieces of a segmented row accumulate into shared logit slots: start them at zero (the kernel zeroes
        // the slots of every other row itself)
        LKG_CUDA(cudaMemsetAsync(values, 0, (size_t)g->nnz * sizeof(float), stream));
    }
    const int total = g->n_relations * dim;
    rel_exp_kernel<<<(total + 255) / 256, 256, 0, stream>>>(relation, ld_relation, g->n_relations, dim, rel_exp);
    LKG_LAUNCH_CHECK("rel_exp_kernel");
    const int nvec = dim / 4;
    const int slots = (nvec + 31) / 32;
    const int block = kWarps * 32;
    static const int ring_env = getenv("LKG_ATTN_RING") ? atoi(getenv("LKG_ATTN_RING")) : 0;
    const int ring = ring_env == 4 ? 4 : 8;
    const size_t smem = attn_smem_bytes(nvec, ring);
    auto launch = [&](auto kern) -> int {
        LKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        LKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem));
        if (per_sm < 1) LKG_FAIL(LKG_ERR_UNSUPPORTED, "attention kernel does not fit (smem %zu)", smem);
        kern<<<sm_count() * per_sm, block, smem, stream>>>(*g, entity, ld_entity, rel_exp, nvec, values, counter);
        return LKG_OK;
    };
    int rc;
    switch (slots) {
        case 1: rc = launch(attn_update_kernel<1, 8, 2>); break;
        case 2: rc = launch(attn_update_kernel<2, 8, 2>); break;
        case 3: rc = ring == 4 ? launch(attn_update_kernel<3, 4, 4>) : launch(attn_update_kernel<3, 8, 2>); break;
        default: rc = launch(attn_update_kernel<4, 8, 1>); break;
    }
    if (rc) return rc;
    LKG_LAUNCH_CHECK("attn_update_kernel");
    return LKG_OK;
}

extern "C" int lkg_attn_run_logits(const int32_t* run_ptr, const int32_t* run_slot, int64_t n_runs, const int32_t* att_tail,
                                   const float* entity, int64_t ld_entity, int32_t dim, const float* run_w, int64_t ld_w,
                                   float* logits, void* stream_) {
    LKG_REQUIRE(run_ptr && run_slot && att_tail && entity && run_w && logits && n_runs >= 0, "null argument");
    LKG_REQUIRE(dim > 0 && dim % 4 == 0 && dim <= 512 && ld_entity % 4 == 0 && ld_w % 4 == 0 && aligned16(entity) &&
                    aligned16(run_w), "projected attention needs dim %% 4 == 0, dim <= 512, 16-byte aligned rows");
    if (n_runs == 0) return LKG_OK;
    const int nvec = dim / 4;
    int64_t blocks = (n_runs * 32 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nvec <= 32)
        attn_run_logits_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(run_ptr, run_slot, n_runs, att_tail, entity, ld_entity,
                                                                        nvec, run_w, ld_w, logits);
    else if (nvec <= 64)
        attn_run_logits_kernel<2><<<(unsigned)blocks, 256, 0, stream>>>(run_ptr, run_slot, n_runs, att_tail, entity, ld_entity,
                                                                        nvec, run_w, ld_w, logits);
    else if (nvec <= 96)
        attn_run_logits_kernel<3><<<(unsigned)blocks, 256, 0, stream>>>(run_ptr, run_slot, n_runs, att_tail, entity, ld_entity,
                                                                        nvec, run_w, ld_w, logits);
    else
        attn_run_logits_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(run_ptr, run_slot, n_runs, att_tail, entity, ld_entity,
                                                                        nvec, run_w, ld_w, logits);
    LKG_LAUNCH_CHECK("attn_run_logits_kernel");
    return LKG_OK;
}

extern "C" int lkg_row_softmax(const int32_t* rowptr, int64_t n_rows, float* values, void* stream_) {
    LKG_REQUIRE(rowptr && values && n_rows >= 0, "null argument");
    if (n_rows == 0) return LKG_OK;
    int64_t blocks = (n_rows * 32 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    row_softmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(rowptr, n_rows, values);
    LKG_LAUNCH_CHECK("row_softmax_kernel");
    return LKG_OK;
}
