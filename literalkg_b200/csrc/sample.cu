// Minibatch assembly on the device: positive / negative sampling of the two training loops.
//   generate_kg_batch           dataloader.py:285-318 (+ sample_pos_triples_for_head :254-271,
//                                                        sample_neg_triples_for_head :273-283)
//   generate_prediction_batch   dataloader.py:221-252 (+ sample_pos_tails_for_head :192-206,
//                                                        sample_neg_tails_for_head :208-219)
// The reference draws, per head, one positive (tail, relation) uniformly from the head's triples and neg_rate
// negative tails from a candidate list by rejection (not a positive of the head under that relation, not drawn
// before), in Python loops over lists -- `random.choice(list(set))` rebuilds a list per draw -- which dominate an
// epoch once the graph pass is fast.  Here one thread owns one head: the head's triples are
// a CSR row of the plan sorted by (relation, tail), so the rejection test is a binary search; the random stream is
// a counter-based generator keyed by (seed, head slot), reproducible for a given seed.
// The head draw itself (random.sample / random.choice of the existing heads) stays on the host side of the binding.
#include "common.cuh"

namespace lkg {
namespace {

// SplitMix64 as a counter-based generator: stream = hash(seed, slot), value i = mix(stream + i * golden)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t state;
    __device__ Rng(uint64_t seed, uint64_t slot) : state(mix64(seed ^ mix64(slot + 0x9e3779b97f4a7c15ull))) {}
    __device__ uint64_t next() { return mix64(state += 0x9e3779b97f4a7c15ull); }
    // unbiased integer in [0, n): 64-bit multiply-high with rejection of the short tail (Lemire)
    __device__ uint64_t below(uint64_t n) {
        uint64_t x = next();
        uint64_t hi = __umul64hi(x, n), lo = x * n;
        if (lo < n) {
            const uint64_t t = (0 - n) % n;
            while (lo < t) {
                x = next();
                hi = __umul64hi(x, n);
                lo = x * n;
            }
        }
        return hi;
    }
};

__global__ void sample_batch_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ tails,
                                    const int32_t* __restrict__ rels, const int64_t* __restrict__ heads, int64_t n_heads,
                                    const int64_t* __restrict__ cand, int64_t n_cand, int neg_rate, int use_relation,
                                    uint64_t seed, int max_tries, int64_t* __restrict__ out_h,
                                    int64_t* __restrict__ out_r, int64_t* __restrict__ out_pos,
                                    int64_t* __restrict__ out_neg, int32_t* __restrict__ n_failed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_heads) return;
    const int64_t h = heads[i];
    const int u0 = __ldg(rowptr + h), u1 = __ldg(rowptr + h + 1);
    Rng rng(seed, (uint64_t)i);
    int64_t pos_t = -1, pos_r = -1;
    if (u1 > u0) {
        const int e = u0 + (int)rng.below((uint64_t)(u1 - u0));
        pos_t = __ldg(tails + e);
        pos_r = __ldg(rels + e);
    } else {
        // a head without triples cannot be sampled (KeyError upstream): flagged, and the ids that are emitted stay
        // inside the tables (the loss kernels do no range checks)
        atomicAdd(n_failed, 1);
        pos_t = cand[0];
        pos_r = 0;
    }
    const int rel = use_relation ? (int)pos_r : -1;
    for (int k = 0; k < neg_rate; ++k) {
        out_h[i * neg_rate + k] = h;
        if (out_r) out_r[i * neg_rate + k] = pos_r;
        out_pos[i * neg_rate + k] = pos_t;
    }
    for (int k = 0; k < neg_rate; ++k) {
        int64_t pick = -1, last = cand[0];
        for (int tries = 0; tries < max_tries && pick < 0; ++tries) {
            const int64_t c = cand[rng.below((uint64_t)n_cand)];
            last = c;
            // binary search for (rel, c) in the sorted row
            int lo = u0, hi = u1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const int r = use_relation ? __ldg(rels + mid) : -1;
                const int t = __ldg(tails + mid);
                if (r < rel || (r == rel && t < (int)c)) lo = mid + 1; else hi = mid;
            }
            bool bad = lo < u1 && (!use_relation || __ldg(rels + lo) == rel) && __ldg(tails + lo) == (int)c;
            for (int q = 0; q < k && !bad; ++q) bad = out_neg[i * neg_rate + q] == c;
            if (!bad) pick = c;
        }
        if (pick < 0) {                              // the reference would loop forever here: flagged (n_failed), and the
            atomicAdd(n_failed, 1);                  // last candidate drawn is emitted so that the id stays valid
            pick = last;
        }
        out_neg[i * neg_rate + k] = pick;
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_sample_batch(const int32_t* rowptr, const int32_t* tails, const int32_t* rels, const int64_t* heads,
                                int64_t n_heads, const int64_t* candidates, int64_t n_candidates, int32_t neg_rate,
                                int32_t use_relation, uint64_t seed, int32_t max_tries, int64_t* out_h, int64_t* out_r,
                                int64_t* out_pos, int64_t* out_neg, int32_t* n_failed, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_heads == 0) return LKG_OK;
    LKG_REQUIRE(rowptr && tails && rels && heads && candidates && out_h && out_pos && out_neg && n_failed,
                "null argument");
    LKG_REQUIRE(n_candidates > 0 && neg_rate > 0 && max_tries > 0, "bad sampling parameters");
    const unsigned blocks = (unsigned)((n_heads + 127) / 128);
    sample_batch_kernel<<<blocks, 128, 0, stream>>>(rowptr, tails, rels, heads, n_heads, candidates, n_candidates,
                                                    neg_rate, use_relation, seed, max_tries, out_h, out_r, out_pos,
                                                    out_neg, n_failed);
    LKG_LAUNCH_CHECK("sample_batch_kernel");
    return LKG_OK;
}
