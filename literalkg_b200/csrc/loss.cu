// Loss heads of the two training modes, forward value and gradients in one pass over the minibatch:
//   lkg_bpr_loss     calculate_prediction_loss  (model.py:316-348)  BPR on dot products of the final embeddings
//   lkg_transr_loss  calc_triplet_loss          (model.py:364-428)  TransR: rows projected by the relation's W_r
// The reference gathers three [B, G] row blocks and a [B, G, D] copy of W_r (209 MB at B = 681), runs three bmm and
// ~15 elementwise / reduction kernels, and autograd replays all of it.  Here one CTA owns one triple: its rows and
// W_r stream through once for the forward and once for the backward (W_r stays in L2: R relations, 20 MB at R = 64),
// and the gradients go straight into d emb (rows of the batch only), d W_r, d relation_embed with atomics.
// Both losses are means over the batch, so every per-sample gradient carries 1 / B.
#include "common.cuh"

namespace lkg {
namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

// -log(sigmoid(x)) = softplus(-x), stable
__device__ __forceinline__ float neg_logsigmoid(float x) { return fmaxf(-x, 0.f) + log1pf(__expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

// ---- BPR: one warp per triple --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bpr_loss_kernel(const float* __restrict__ emb, int64_t ld, int g_dim,
                                                       const int64_t* __restrict__ h, const int64_t* __restrict__ pos,
                                                       const int64_t* __restrict__ neg, int64_t batch, float lambda,
                                                       float* __restrict__ loss, const float* __restrict__ grad_scale,
                                                       float* __restrict__ d_emb, int64_t ld_d) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= batch) return;
    const float* hr = emb + h[i] * ld;
    const float* pr = emb + pos[i] * ld;
    const float* nr = emb + neg[i] * ld;
    float sp = 0.f, sn = 0.f, qh = 0.f, qp = 0.f, qn = 0.f;
    for (int c = lane; c < g_dim; c += 32) {
        const float a = __ldg(hr + c), b = __ldg(pr + c), d = __ldg(nr + c);
        sp = fmaf(a, b, sp);
        sn = fmaf(a, d, sn);
        qh = fmaf(a, a, qh);
        qp = fmaf(b, b, qp);
        qn = fmaf(d, d, qn);
    }
    sp = warp_sum(sp); sn = warp_sum(sn); qh = warp_sum(qh); qp = warp_sum(qp); qn = warp_sum(qn);
    const float inv_b = 1.f / (float)batch;
    const float x = sp - sn;
    if (loss && lane == 0) atomicAdd(loss, (neg_logsigmoid(x) + lambda * 0.5f * (qh + qp + qn)) * inv_b);
    if (d_emb) {
        const float up = grad_scale ? __ldg(grad_scale) : 1.f;   // upstream d / d loss
        const float s = -sigmoid_f(-x) * inv_b * up;     // d loss / d (pos_score - neg_score)
        const float l = lambda * inv_b * up;
        float* dh = d_emb + h[i] * ld_d;
        float* dp = d_emb + pos[i] * ld_d;
        float* dn = d_emb + neg[i] * ld_d;
        for (int c = lane; c < g_dim; c += 32) {
            const float a = __ldg(hr + c), b = __ldg(pr + c), d = __ldg(nr + c);
            atomicAdd(dh + c, fmaf(s, b - d, l * a));
            atomicAdd(dp + c, fmaf(s, a, l * b));
            atomicAdd(dn + c, fmaf(-s, a, l * d));
        }
    }
}

// ---- TransR: one CTA per triple --------------------------------------------------------------------------------
struct TransRParams {
    const float* emb;   int64_t ld;
    int g_dim, r_dim;
    const float* rel;   int64_t ld_rel;       // relation_embed [R, r_dim]
    const float* m;                           // gat_trans_M [R, g_dim, r_dim]
    const int64_t* h; const int64_t* r; const int64_t* pos; const int64_t* neg;
    int64_t batch;
    float lambda;
    float* loss;                              // nullable (backward only)
    const float* grad_scale;                  // nullable device scalar: upstream d / d loss
    float* d_emb;       int64_t ld_d;         // nullable (forward only)
    float* d_rel;                             // [R, r_dim]
    float* d_m;                               // [R, g_dim, r_dim]
};

__global__ void __launch_bounds__(256) transr_loss_kernel(TransRParams p) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[8];
    const int G = p.g_dim, D = p.r_dim;
    float* sh = sm;                 // [G] head row
    float* sp = sh + G;             // [G] positive tail
    float* sn = sp + G;             // [G] negative tail
    float* sa = sn + G;             // [D] h W      -> d a
    float* sb = sa + D;             // [D] pos W    -> d b
    float* sc = sb + D;             // [D] neg W    -> d c
    const int64_t i = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rid = p.r[i];
    const float* W = p.m + rid * (int64_t)G * D;
    const float* er = p.rel + rid * p.ld_rel;
    const int64_t ih = p.h[i], ip = p.pos[i], in_ = p.neg[i];
    for (int g = tid; g < G; g += blockDim.x) {
        sh[g] = __ldg(p.emb + ih * p.ld + g);
        sp[g] = __ldg(p.emb + ip * p.ld + g);
        sn[g] = __ldg(p.emb + in_ * p.ld + g);
    }
    __syncthreads();
    // a = h W, b = pos W, c = neg W: thread j owns output column j, W rows read coalesced
    float pos_s = 0.f, neg_s = 0.f, qa = 0.f, qb = 0.f, qc = 0.f, qe = 0.f;
    for (int j = tid; j < D; j += blockDim.x) {
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll 4
        for (int g = 0; g < G; ++g) {
            const float w = __ldg(W + (int64_t)g * D + j);
            a = fmaf(sh[g], w, a);
            b = fmaf(sp[g], w, b);
            c = fmaf(sn[g], w, c);
        }
        const float e = __ldg(er + j);
        const float u = a + e - b, v = a + e - c;
        pos_s = fmaf(u, u, pos_s);
        neg_s = fmaf(v, v, neg_s);
        qa = fmaf(a, a, qa); qb = fmaf(b, b, qb); qc = fmaf(c, c, qc); qe = fmaf(e, e, qe);
        sa[j] = a; sb[j] = b; sc[j] = c;
    }
    pos_s = block_sum(pos_s, red);
    neg_s = block_sum(neg_s, red);
    const float l2 = block_sum(qa + qb + qc + qe, red);
    const float inv_b = 1.f / (float)p.batch;
    const float x = neg_s - pos_s;
    if (p.loss && tid == 0) atomicAdd(p.loss, (neg_logsigmoid(x) + p.lambda * 0.5f * l2) * inv_b);
    if (!p.d_emb) return;

    // d a, d b, d c, d e_r in place
    const float up = p.grad_scale ? __ldg(p.grad_scale) : 1.f;
    const float s = -sigmoid_f(-x) * inv_b * up;         // d loss / d (neg_score - pos_score)
    const float l = p.lambda * inv_b * up;
    __syncthreads();
    for (int j = tid; j < D; j += blockDim.x) {
        const float a = sa[j], b = sb[j], c = sc[j], e = __ldg(er + j);
        const float u = a + e - b, v = a + e - c;
        const float common = 2.f * s * (v - u);          // d/d a and d/d e_r of s (|v|^2 - |u|^2)
        sa[j] = common + l * a;
        sb[j] = 2.f * s * u + l * b;
        sc[j] = -2.f * s * v + l * c;
        atomicAdd(p.d_rel + rid * D + j, common + l * e);
    }
    __syncthreads();
    // d rows = W d(.) ; d W += row (x) d(.): one warp per W row, lanes over the r_dim columns
    float* dW = p.d_m + rid * (int64_t)G * D;
    for (int g = warp; g < G; g += (blockDim.x >> 5)) {
        const float hg = sh[g], pg = sp[g], ng = sn[g];
        float dh = 0.f, dp = 0.f, dn = 0.f;
        for (int j = lane; j < D; j += 32) {
            const float w = __ldg(W + (int64_t)g * D + j);
            const float da = sa[j], db = sb[j], dc = sc[j];
            dh = fmaf(w, da, dh);
            dp = fmaf(w, db, dp);
            dn = fmaf(w, dc, dn);
            atomicAdd(dW + (int64_t)g * D + j, fmaf(hg, da, fmaf(pg, db, ng * dc)));
        }
        dh = warp_sum(dh); dp = warp_sum(dp); dn = warp_sum(dn);
        if (lane == 0) {
            atomicAdd(p.d_emb + ih * p.ld_d + g, dh);
            atomicAdd(p.d_emb + ip * p.ld_d + g, dp);
            atomicAdd(p.d_emb + in_ * p.ld_d + g, dn);
        }
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_bpr_loss(const float* emb, int64_t ld_emb, int32_t g_dim, const int64_t* h, const int64_t* pos,
                            const int64_t* neg, int64_t batch, float l2_lambda, float* loss, const float* grad_scale,
                            float* d_emb, int64_t ld_d, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(emb && h && pos && neg && (loss || d_emb) && g_dim > 0 && batch > 0, "bad BPR arguments");
    const int64_t blocks = (batch * 32 + 255) / 256;
    bpr_loss_kernel<<<(unsigned)blocks, 256, 0, stream>>>(emb, ld_emb, g_dim, h, pos, neg, batch, l2_lambda, loss,
                                                         grad_scale, d_emb, ld_d);
    LKG_LAUNCH_CHECK("bpr_loss_kernel");
    return LKG_OK;
}

extern "C" int lkg_transr_loss(const float* emb, int64_t ld_emb, int32_t g_dim, const float* relation, int64_t ld_rel,
                               int32_t r_dim, const float* trans_m, const int64_t* h, const int64_t* r,
                               const int64_t* pos, const int64_t* neg, int64_t batch, float l2_lambda, float* loss,
                               const float* grad_scale, float* d_emb, int64_t ld_d, float* d_relation,
                               float* d_trans_m, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(emb && relation && trans_m && h && r && pos && neg && (loss || d_emb) && g_dim > 0 && r_dim > 0 &&
                    batch > 0, "bad TransR arguments");
    LKG_REQUIRE(!d_emb || (d_relation && d_trans_m), "the backward needs all three gradient buffers");
    TransRParams p{emb, ld_emb, g_dim, r_dim, relation, ld_rel, trans_m, h, r, pos, neg, batch, l2_lambda, loss,
                   grad_scale, d_emb, ld_d, d_relation, d_trans_m};
    const size_t smem = (size_t)(3 * g_dim + 3 * r_dim) * sizeof(float);
    if (smem > 200 * 1024) LKG_FAIL(LKG_ERR_UNSUPPORTED, "TransR loss: dims %d / %d do not fit shared memory", g_dim, r_dim);
    LKG_CUDA(cudaFuncSetAttribute(transr_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    transr_loss_kernel<<<(unsigned)batch, 256, smem, stream>>>(p);
    LKG_LAUNCH_CHECK("transr_loss_kernel");
    return LKG_OK;
}
