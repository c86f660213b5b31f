// Graph plan: CSR construction and the initial Laplacian attention matrix.
//
// Replaces the tensor-producing part of DataLoader.construct_data (dataloader.py:369-424), the
// coalesce + sort hidden inside torch.sparse.softmax (model.py:466-470) and
// create_adjacency_dict / create_laplacian_dict (dataloader.py:449-495).
//
// Integer work, HBM bound.  Sorting uses CUB's device radix sort (part of the CUDA toolkit); every
// other step is a hand-written grid-stride kernel.  All outputs are bit-exact w.r.t. the reference:
// the (h,t) pair list equals A_in.coalesce().indices().
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace lkg {
namespace {

constexpr uint64_t kDropped = ~0ull;

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct PlanScratch {
    uint64_t* key_a;   // [E]
    uint64_t* key_b;   // [E]
    int32_t* val_a;    // [E]
    int32_t* val_b;    // [E]
    uint64_t* key_c;   // [E]
    int32_t* val_c;    // [E]
    int32_t* seg_ht;   // [E] segment id of each triple in (h,t) order
    int32_t* deg_e;    // [N+1]
    int32_t* deg_u;    // [N+1]
    int32_t* ord_key;  // [N] sorted degrees (discarded)
    int32_t* ord_val;  // [N] row ids 0..N-1
    void* cub;         // cub temp storage
    size_t cub_bytes;
    size_t total;
};

size_t cub_temp_bytes(int64_t e, int64_t n) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t*)nullptr, (uint64_t*)nullptr, (int32_t*)nullptr,
                                    (int32_t*)nullptr, (int)e, 0, 64);
    cub::DeviceScan::InclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, (int)e);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (int32_t*)nullptr, (int32_t*)nullptr, (int)(n + 1));
    size_t d = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, d, (int32_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr,
                                              (int32_t*)nullptr, (int)n, 0, 32);
    size_t m = a > b ? a : b;
    m = m > c ? m : c;
    return m > d ? m : d;
}

PlanScratch carve(void* base, int64_t e, int64_t n) {
    PlanScratch s{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes);
        return p;
    };
    const size_t ee = (size_t)(e > 0 ? e : 1);
    s.key_a = (uint64_t*)take(ee * 8);
    s.key_b = (uint64_t*)take(ee * 8);
    s.val_a = (int32_t*)take(ee * 4);
    s.val_b = (int32_t*)take(ee * 4);
    s.key_c = (uint64_t*)take(ee * 8);
    s.val_c = (int32_t*)take(ee * 4);
    s.seg_ht = (int32_t*)take(ee * 4);
    s.deg_e = (int32_t*)take((size_t)(n + 1) * 4);
    s.deg_u = (int32_t*)take((size_t)(n + 1) * 4);
    s.ord_key = (int32_t*)take((size_t)n * 4);
    s.ord_val = (int32_t*)take((size_t)n * 4);
    s.cub_bytes = cub_temp_bytes(e, n);
    s.cub = take(s.cub_bytes);
    s.total = off;
    return s;
}

// key = (h << 32) | t, payload = position in the input.  Dropped / invalid triples get the all-ones
// key and sort last.
__global__ void make_ht_keys(const int64_t* __restrict__ h, const int64_t* __restrict__ t,
                             const int64_t* __restrict__ r, int64_t e, int64_t n, int32_t n_rel,
                             const uint8_t* __restrict__ keep, uint64_t* __restrict__ key,
                             int32_t* __restrict__ val, int64_t* __restrict__ counts) {
    int64_t dropped = 0, bad = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hh = h[i], tt = t[i], rr = r[i];
        const bool ok = hh >= 0 && hh < n && tt >= 0 && tt < n && rr >= 0 && rr < n_rel;
        const bool kept = ok && (keep == nullptr || keep[rr] != 0);
        key[i] = kept ? ((uint64_t)hh << 32) | (uint64_t)tt : kDropped;
        val[i] = (int32_t)i;
        dropped += kept ? 0 : 1;
        bad += ok ? 0 : 1;
    }
    // one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        dropped += __shfl_xor_sync(kFull, dropped, o);
        bad += __shfl_xor_sync(kFull, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (dropped) atomicAdd((unsigned long long*)&counts[0], (unsigned long long)dropped);
        if (bad) atomicAdd((unsigned long long*)&counts[2], (unsigned long long)bad);
    }
}

// counts[0] currently holds #dropped -> turn into #kept.
__global__ void finalize_kept(int64_t e, int64_t* counts) { counts[0] = e - counts[0]; }

// flag[i] = 1 where a new (h,t) pair starts (input sorted by (h,t)); also per-row triple degree.
__global__ void mark_pairs(const uint64_t* __restrict__ key, const int64_t* __restrict__ counts,
                           int32_t* __restrict__ flag, int32_t* __restrict__ deg_e) {
    const int64_t e = counts[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = key[i];
        flag[i] = (i == 0 || key[i - 1] != k) ? 1 : 0;
        atomicAdd(&deg_e[k >> 32], 1);
    }
}

// seg = inclusive_scan(flag) - 1.  Emits the unique pair list, per-row unique degree and the
// (h, r) keys for the second (stable) sort with payload = position in (h,t) order.
__global__ void emit_pairs(const uint64_t* __restrict__ key, const int32_t* __restrict__ src_pos,
                           const int64_t* __restrict__ rel_in, int32_t* __restrict__ file_seg,
                           int32_t* __restrict__ seg_inclusive, const int64_t* __restrict__ counts_in,
                           int64_t* __restrict__ counts, int32_t* __restrict__ col,
                           int64_t* __restrict__ coo_rows, int64_t* __restrict__ coo_cols,
                           int32_t* __restrict__ deg_u, uint64_t* __restrict__ key_hr,
                           int32_t* __restrict__ pos, int64_t e_total) {
    const int64_t e = counts_in[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e_total; i += (int64_t)gridDim.x * blockDim.x) {
        if (i >= e) {               // dropped triples must stay last in the second sort as well
            key_hr[i] = kDropped;
            pos[i] = (int32_t)i;
            if (file_seg) file_seg[src_pos[i]] = -1;
            continue;
        }
        const uint64_t k = key[i];
        const int32_t s = seg_inclusive[i] - 1;
        const bool first = (i == 0) || (key[i - 1] != k);
        seg_inclusive[i] = s;
        if (first) {
            const int32_t hh = (int32_t)(k >> 32), tt = (int32_t)(k & 0xffffffffu);
            col[s] = tt;
            if (coo_rows) coo_rows[s] = hh;
            if (coo_cols) coo_cols[s] = tt;
            atomicAdd(&deg_u[hh], 1);
        }
        if (i == e - 1) counts[1] = (int64_t)s + 1;
        if (file_seg) file_seg[src_pos[i]] = s;
        key_hr[i] = (k & 0xffffffff00000000ull) | (uint32_t)rel_in[src_pos[i]];
        pos[i] = (int32_t)i;
    }
    if (e == 0 && blockIdx.x == 0 && threadIdx.x == 0) counts[1] = 0;
}

// gather the att-order arrays through the permutation produced by the (h, r) sort
__global__ void gather_att(const uint64_t* __restrict__ key_hr_sorted, const int32_t* __restrict__ pos_sorted,
                           const uint64_t* __restrict__ key_ht, const int32_t* __restrict__ seg_ht,
                           const int64_t* __restrict__ counts, int32_t* __restrict__ att_tail,
                           int32_t* __restrict__ att_rel, int32_t* __restrict__ att_seg) {
    const int64_t e = counts[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = pos_sorted[i];
        const uint64_t k = key_ht[p];
        att_tail[i] = (int32_t)(k & 0xffffffffu);
        att_rel[i] = (int32_t)(key_hr_sorted[i] & 0xffffffffu);
        // bit 31 flags a triple whose (h,t) pair has other triples (another relation): its logit must be ADDED to the
        // pair slot; every other triple owns its slot and is simply stored
        const bool dup = (p > 0 && key_ht[p - 1] == k) || (p + 1 < e && key_ht[p + 1] == k);
        att_seg[i] = seg_ht[p] | (dup ? (int32_t)0x80000000 : 0);
    }
}

__global__ void iota_kernel(int32_t* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int32_t)i;
}

// schedule record of the i-th heaviest row: {row, att_begin, att_end, agg_begin, agg_end, 0, 0, 0} -- one 32-byte
// sector tells a warp everything about its next row (instead of a row_order -> rowptr -> rowptr chain of loads)
__global__ void make_sched_kernel(const int32_t* __restrict__ row_order, const int32_t* __restrict__ att_rowptr,
                                  const int32_t* __restrict__ rowptr, int64_t n, int32_t* __restrict__ sched) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t row = row_order[i];
        int4* out = reinterpret_cast<int4*>(sched + 8 * i);
        out[0] = make_int4(row, att_rowptr[row], att_rowptr[row + 1], rowptr[row]);
        out[1] = make_int4(rowptr[row + 1], 0, 0, 0);
    }
}

inline int grid_for(int64_t n, int block = 256) {
    int64_t b = (n + block - 1) / block;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- Laplacian -------------------------------------------------------------------------------
constexpr int32_t kSegMask = 0x7fffffff;   // att_seg: bit 31 = "pair has several triples"

// one thread per head row; walks the row's (h,r) runs in att order.  deg_r(h) = run length.
// random-walk: each triple contributes 1/deg_r(h); symmetric: deg_r(h)^-1/2 * deg_r(t)^-1/2 where
// deg_r(t) is the OUT-degree of t under r (row sums on both sides, dataloader.py:464-470), 0 -> 0.
__device__ int32_t run_length(const lkg_graph& g, int32_t node, int32_t rel) {
    int32_t lo = g.att_rowptr[node], hi = g.att_rowptr[node + 1];
    const int32_t end = hi;
    // lower bound of rel
    int32_t a = lo, b = hi;
    while (a < b) {
        const int32_t m = (a + b) >> 1;
        if (g.att_rel[m] < rel) a = m + 1; else b = m;
    }
    const int32_t first = a;
    b = end;
    while (a < b) {
        const int32_t m = (a + b) >> 1;
        if (g.att_rel[m] <= rel) a = m + 1; else b = m;
    }
    return a - first;
}

__global__ void laplacian_kernel(lkg_graph g, int symmetric, double* __restrict__ acc) {
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < g.n_entities;
         row += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e0 = g.att_rowptr[row], e1 = g.att_rowptr[row + 1];
        int32_t i = e0;
        while (i < e1) {
            const int32_t rel = g.att_rel[i];
            int32_t j = i + 1;
            while (j < e1 && g.att_rel[j] == rel) ++j;
            const double deg = (double)(j - i);
            const double dh = symmetric ? 1.0 / sqrt(deg) : 1.0 / deg;
            for (int32_t k = i; k < j; ++k) {
                double v = dh;
                if (symmetric) {
                    const int32_t dt = run_length(g, g.att_tail[k], rel);
                    v = dt > 0 ? dh * (1.0 / sqrt((double)dt)) : 0.0;
                }
                acc[g.att_seg[k] & kSegMask] += v;   // all triples of one pair live in this row: no race
            }
            i = j;
        }
    }
}

__global__ void cast_f64_f32(const double* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_plan_workspace_bytes(int64_t n_edges, int64_t n_entities, size_t* bytes) {
    LKG_REQUIRE(bytes != nullptr, "bytes is null");
    LKG_REQUIRE(n_edges >= 0 && n_edges < (1ll << 31) - 1, "n_edges out of range");
    LKG_REQUIRE(n_entities > 0 && n_entities < (1ll << 31) - 1, "n_entities out of range");
    *bytes = carve(nullptr, n_edges, n_entities).total;
    return LKG_OK;
}

extern "C" int lkg_plan_build(const int64_t* h, const int64_t* t, const int64_t* r, int64_t n_edges,
                              int64_t n_entities, int32_t n_relations, const uint8_t* rel_keep,
                              int32_t* att_rowptr, int32_t* att_tail, int32_t* att_rel, int32_t* att_seg,
                              int32_t* rowptr, int32_t* col, int32_t* row_order, int32_t* row_sched, int64_t* coo_rows,
                              int64_t* coo_cols, int32_t* file_seg, int64_t* counts_dev, void* workspace,
                              size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(n_edges >= 0 && n_edges < (1ll << 31) - 1, "n_edges out of range");
    LKG_REQUIRE(n_entities > 0 && n_entities < (1ll << 31) - 1, "n_entities out of range");
    LKG_REQUIRE(n_relations > 0, "n_relations must be positive");
    LKG_REQUIRE(att_rowptr && rowptr && counts_dev, "null output");
    LKG_REQUIRE(n_edges == 0 || (h && t && r && att_tail && att_rel && att_seg && col), "null edge array");
    PlanScratch s = carve(workspace, n_edges, n_entities);
    if (workspace == nullptr || workspace_bytes < s.total)
        LKG_FAIL(LKG_ERR_WORKSPACE, "plan workspace too small: %zu < %zu", workspace_bytes, s.total);

    const int64_t n1 = n_entities + 1;
    LKG_CUDA(cudaMemsetAsync(counts_dev, 0, 3 * sizeof(int64_t), stream));
    LKG_CUDA(cudaMemsetAsync(s.deg_e, 0, n1 * 4, stream));
    LKG_CUDA(cudaMemsetAsync(s.deg_u, 0, n1 * 4, stream));
    const int e32 = (int)n_edges;
    if (n_edges > 0) {
        make_ht_keys<<<grid_for(n_edges), 256, 0, stream>>>(h, t, r, n_edges, n_entities, n_relations,
                                                            rel_keep, s.key_a, s.val_a, counts_dev);
        LKG_LAUNCH_CHECK("make_ht_keys");
    }
    finalize_kept<<<1, 1, 0, stream>>>(n_edges, counts_dev);
    LKG_LAUNCH_CHECK("finalize_kept");
    if (n_edges > 0) {
        size_t tb = s.cub_bytes;
        // pass 1: (h,t) order, payload = input position.  Dropped triples carry the all-ones key -> last.
        LKG_CUDA(cub::DeviceRadixSort::SortPairs(s.cub, tb, s.key_a, s.key_b, s.val_a, s.val_b, e32, 0, 64, stream));
        mark_pairs<<<grid_for(n_edges), 256, 0, stream>>>(s.key_b, counts_dev, s.seg_ht, s.deg_e);
        LKG_LAUNCH_CHECK("mark_pairs");
        // entries past E-kept hold garbage flags; they are never read back (all consumers stop at E-kept)
        tb = s.cub_bytes;
        LKG_CUDA(cub::DeviceScan::InclusiveSum(s.cub, tb, s.seg_ht, s.seg_ht, e32, stream));
        emit_pairs<<<grid_for(n_edges), 256, 0, stream>>>(s.key_b, s.val_b, r, file_seg, s.seg_ht, counts_dev, counts_dev, col,
                                                          coo_rows, coo_cols, s.deg_u, s.key_a, s.val_a, n_edges);
        LKG_LAUNCH_CHECK("emit_pairs");
        // pass 2: stable sort by (h, r) of the (h,t)-ordered list -> (h, r, t) order
        tb = s.cub_bytes;
        LKG_CUDA(cub::DeviceRadixSort::SortPairs(s.cub, tb, s.key_a, s.key_c, s.val_a, s.val_c, e32, 0, 64, stream));
        gather_att<<<grid_for(n_edges), 256, 0, stream>>>(s.key_c, s.val_c, s.key_b, s.seg_ht, counts_dev,
                                                          att_tail, att_rel, att_seg);
        LKG_LAUNCH_CHECK("gather_att");
    }
    size_t tb = s.cub_bytes;
    if (row_order) {
        // rows by decreasing triple count; the LSD radix sort is stable, so equal degrees keep ascending row ids
        iota_kernel<<<grid_for(n_entities), 256, 0, stream>>>(s.ord_val, n_entities);
        LKG_LAUNCH_CHECK("iota_kernel");
        LKG_CUDA(cub::DeviceRadixSort::SortPairsDescending(s.cub, tb, s.deg_e, s.ord_key, s.ord_val, row_order,
                                                           (int)n_entities, 0, 32, stream));
        tb = s.cub_bytes;
    }
    LKG_CUDA(cub::DeviceScan::ExclusiveSum(s.cub, tb, s.deg_e, att_rowptr, (int)n1, stream));
    tb = s.cub_bytes;
    LKG_CUDA(cub::DeviceScan::ExclusiveSum(s.cub, tb, s.deg_u, rowptr, (int)n1, stream));
    if (row_sched) {
        LKG_REQUIRE(row_order != nullptr && aligned16(row_sched), "row_sched needs row_order and 16-byte alignment");
        make_sched_kernel<<<grid_for(n_entities), 256, 0, stream>>>(row_order, att_rowptr, rowptr, n_entities, row_sched);
        LKG_LAUNCH_CHECK("make_sched_kernel");
    }
    return LKG_OK;
}

extern "C" int lkg_laplacian_init(const lkg_graph* g, int symmetric, float* values, double* scratch,
                                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(g != nullptr, "graph is null");
    LKG_REQUIRE(g->nnz == 0 || (values != nullptr && scratch != nullptr), "values / scratch is null");
    if (g->nnz == 0) return LKG_OK;
    LKG_CUDA(cudaMemsetAsync(scratch, 0, (size_t)g->nnz * sizeof(double), stream));
    laplacian_kernel<<<grid_for(g->n_entities, 128), 128, 0, stream>>>(*g, symmetric, scratch);
    LKG_LAUNCH_CHECK("laplacian_kernel");
    cast_f64_f32<<<grid_for(g->nnz), 256, 0, stream>>>(scratch, values, g->nnz);
    LKG_LAUNCH_CHECK("cast_f64_f32");
    return LKG_OK;
}

namespace lkg {
namespace {
__global__ void scatter_add_kernel(const float* __restrict__ in, const int32_t* __restrict__ seg, int64_t n,
                                   float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = seg[i];
        if (s >= 0) atomicAdd(out + s, in[i]);
    }
}
}  // namespace
}  // namespace lkg

extern "C" int lkg_segment_scatter_add(const float* values_in, const int32_t* file_seg, int64_t n_edges,
                                       float* values_out, int64_t nnz, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(n_edges == 0 || (values_in && file_seg), "null input");
    LKG_REQUIRE(nnz == 0 || values_out, "null output");
    if (nnz > 0) LKG_CUDA(cudaMemsetAsync(values_out, 0, (size_t)nnz * sizeof(float), stream));
    if (n_edges == 0) return LKG_OK;
    scatter_add_kernel<<<grid_for(n_edges), 256, 0, stream>>>(values_in, file_seg, n_edges, values_out);
    LKG_LAUNCH_CHECK("scatter_add_kernel");
    return LKG_OK;
}

namespace lkg {
namespace {
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__global__ void fingerprint_kernel(const int64_t* __restrict__ h, const int64_t* __restrict__ t,
                                   const int64_t* __restrict__ r, int64_t n, unsigned long long* __restrict__ fp) {
    uint64_t a = 0, b = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t x = mix64((uint64_t)h[i] * 0x9e3779b97f4a7c15ull + (uint64_t)t[i]) ^ mix64((uint64_t)r[i] + 0x165667b19e3779f9ull * (uint64_t)t[i]);
        a += x;                       // commutative: independent of the edge order
        b += mix64(x);
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(kFull, a, o);
        b += __shfl_xor_sync(kFull, b, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(fp, (unsigned long long)a);
        atomicAdd(fp + 1, (unsigned long long)b);
    }
}
}  // namespace
}  // namespace lkg

extern "C" int lkg_edge_fingerprint(const int64_t* h, const int64_t* t, const int64_t* r, int64_t n_edges,
                                    uint64_t* fp_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(fp_dev != nullptr, "fp_dev is null");
    LKG_REQUIRE(n_edges == 0 || (h && t && r), "null edge array");
    LKG_CUDA(cudaMemsetAsync(fp_dev, 0, 2 * sizeof(uint64_t), stream));
    if (n_edges == 0) return LKG_OK;
    fingerprint_kernel<<<grid_for(n_edges), 256, 0, stream>>>(h, t, r, n_edges, (unsigned long long*)fp_dev);
    LKG_LAUNCH_CHECK("fingerprint_kernel");
    return LKG_OK;
}
