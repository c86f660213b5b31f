// Variant heads of the reference:
//   lkg_transe_loss                 model_bce.py:329-368   TransE triplet loss on rows of the final embeddings
//   lkg_mlp_fc_fwd / lkg_bn_finalize / lkg_mlp_fc_bwd_weight / lkg_mlp_fc_bwd_input / lkg_bn_relu_bwd / lkg_sigmoid_bwd
//                                   model.py:499-519, model_bce.py:423-436   the `mlp` mode head
//       x = [emb[h] | emb[t]] -> BatchNorm(relu(fc1 x)) -> BatchNorm(relu(fc2 .)) -> sigmoid(fc3 .)
// The head works on one minibatch (<= a few thousand pairs, widths 2G -> 128 -> 64 -> 1): ~0.15 GFLOP, far below a
// tensor-core tile's worth of work, and BatchNorm needs the whole batch between layers.  It is therefore a short
// chain of fp32 SIMT kernels, each fused with what surrounds its GEMM:
//   fc forward   = input transform (row-pair gather of the embedding matrix, or the folded BatchNorm affine of the
//                  previous layer) + GEMM + bias + ReLU / sigmoid + the column sums the next BatchNorm needs;
//   fc backward  = the same input transform recomputed on the fly (no [B, 2G] copy of the gathered rows and no copy of
//                  a BatchNorm output is ever written), dW / db reduced over the batch, dX either scattered straight
//                  into d emb rows or handed to the fused BatchNorm + ReLU backward with its two column sums.
#include "common.cuh"

namespace lkg {
namespace {

constexpr int kTM = 32;      // batch rows per CTA
constexpr int kTN = 64;      // output columns per CTA
constexpr int kTK = 32;      // K chunk
constexpr int kFcThreads = 256;

// Element (b, k) of the layer input.  pair mode: k < half -> src[ia[b], k], else src[ib[b], k - half];
// affine mode: src[b, k] * scale[k] + shift[k] (BatchNorm folded into scale / shift); plain: src[b, k].
struct FcInput {
    const float* src;
    int64_t ld;
    const int64_t* ia;
    const int64_t* ib;
    int half;
    const float* scale;
    const float* shift;
};
__device__ __forceinline__ float fc_in(const FcInput& in, int64_t b, int k) {
    if (in.ia) {
        const int64_t row = k < in.half ? in.ia[b] : in.ib[b];
        return __ldg(in.src + row * in.ld + (k < in.half ? k : k - in.half));
    }
    const float v = __ldg(in.src + b * in.ld + k);
    return in.scale ? fmaf(v, __ldg(in.scale + k), __ldg(in.shift + k)) : v;
}

enum { kActNone = 0, kActRelu = 1, kActSigmoid = 2 };

// out[b, j] = act(sum_k in(b, k) W[j, k] + bias[j]);  stats (nullable, double[2 n]) += column sums of out and out^2
__global__ void __launch_bounds__(kFcThreads) fc_fwd_kernel(FcInput in, int64_t m, int k, const float* __restrict__ w,
                                                            int64_t ldw, const float* __restrict__ bias, int n, int act,
                                                            float* __restrict__ out, int64_t ldo,
                                                            double* __restrict__ stats) {
    __shared__ float xs[kTM][kTK + 1];
    __shared__ float ws[kTN][kTK + 1];
    __shared__ float cs[2][kTN];
    const int tid = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * kTM;
    const int j0 = blockIdx.y * kTN;
    // thread -> 2 rows x 4 columns: rows r0, r0 + 16; columns c0 + {0, 16, 32, 48}
    const int r0 = tid >> 4, c0 = tid & 15;
    float acc[2][4] = {};
    for (int kk = 0; kk < k; kk += kTK) {
        for (int i = tid; i < kTM * kTK; i += kFcThreads) {
            const int r = i / kTK, c = i % kTK;
            xs[r][c] = (b0 + r < m && kk + c < k) ? fc_in(in, b0 + r, kk + c) : 0.f;
        }
        for (int i = tid; i < kTN * kTK; i += kFcThreads) {
            const int r = i / kTK, c = i % kTK;
            ws[r][c] = (j0 + r < n && kk + c < k) ? __ldg(w + (int64_t)(j0 + r) * ldw + kk + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < kTK; ++c) {
            const float x0 = xs[r0][c], x1 = xs[r0 + 16][c];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float wv = ws[c0 + 16 * q][c];
                acc[0][q] = fmaf(x0, wv, acc[0][q]);
                acc[1][q] = fmaf(x1, wv, acc[1][q]);
            }
        }
        __syncthreads();
    }
    if (tid < 2 * kTN) cs[tid / kTN][tid % kTN] = 0.f;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int j = j0 + c0 + 16 * q;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int64_t b = b0 + r0 + 16 * rr;
            if (b < m && j < n) {
                float v = acc[rr][q] + (bias ? __ldg(bias + j) : 0.f);
                if (act == kActRelu) v = fmaxf(v, 0.f);
                if (act == kActSigmoid) v = 1.f / (1.f + __expf(-v));
                out[b * ldo + j] = v;
                s1 += v;
                s2 = fmaf(v, v, s2);
            }
        }
        if (stats && j < n) {
            atomicAdd(&cs[0][c0 + 16 * q], s1);
            atomicAdd(&cs[1][c0 + 16 * q], s2);
        }
    }
    if (stats) {
        __syncthreads();
        if (tid < kTN && j0 + tid < n) {
            atomicAdd(stats + j0 + tid, (double)cs[0][tid]);
            atomicAdd(stats + n + j0 + tid, (double)cs[1][tid]);
        }
    }
}

// BatchNorm1d bookkeeping (torch.nn.BatchNorm1d: biased variance to normalise, unbiased for the running estimate).
// training: mean / var from stats; eval: from the running buffers.  Writes the folded affine scale = gamma * rstd,
// shift = beta - mean * scale and the (mean, rstd) the backward needs.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int64_t m, int n, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int training,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float mean, var;
    if (training) {
        const double mu = stats[j] / (double)m;
        double v = stats[n + j] / (double)m - mu * mu;
        v = v > 0.0 ? v : 0.0;
        mean = (float)mu;
        var = (float)v;
        if (running_mean) {
            const float unbiased = m > 1 ? (float)(v * (double)m / (double)(m - 1)) : var;
            running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * mean;
            running_var[j] = (1.f - momentum) * running_var[j] + momentum * unbiased;
        }
    } else {
        mean = running_mean[j];
        var = running_var[j];
    }
    const float rstd = rsqrtf(var + eps);
    const float sc = __ldg(gamma + j) * rstd;
    scale[j] = sc;
    shift[j] = __ldg(beta + j) - mean * sc;
    if (mean_out) mean_out[j] = mean;
    if (rstd_out) rstd_out[j] = rstd;
}

// dW[j, k] += sum_b dz[b, j] in(b, k);  db[j] += sum_b dz[b, j].  Grid: (k tiles, j tiles, batch splits).
__global__ void __launch_bounds__(kFcThreads) fc_bwd_weight_kernel(const float* __restrict__ dz, int64_t ld_dz, FcInput in,
                                                                   int64_t m, int k, int n, float* __restrict__ dw,
                                                                   int64_t ld_dw, float* __restrict__ db) {
    __shared__ float zs[kTK][kTN + 1];     // [batch chunk][j]
    __shared__ float xs[kTK][kTM + 1];     // [batch chunk][k]
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * kTM, j0 = blockIdx.y * kTN;
    const int64_t per = (m + gridDim.z - 1) / gridDim.z;
    const int64_t bb = blockIdx.z * per, be = bb + per < m ? bb + per : m;
    const int kk = tid & 31, jq = tid >> 5;                  // thread -> k column kk, j columns jq + 8 q
    float acc[8] = {};
    float bsum = 0.f;                                        // threads 0..kTN-1 of the k0 == 0 CTAs: bias gradient
    for (int64_t b = bb; b < be; b += kTK) {
        for (int i = tid; i < kTK * kTN; i += kFcThreads) {
            const int r = i / kTN, c = i % kTN;
            zs[r][c] = (b + r < be && j0 + c < n) ? __ldg(dz + (b + r) * ld_dz + j0 + c) : 0.f;
        }
        for (int i = tid; i < kTK * kTM; i += kFcThreads) {
            const int r = i / kTM, c = i % kTM;
            xs[r][c] = (b + r < be && k0 + c < k) ? fc_in(in, b + r, k0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < kTK; ++r) {
            const float x = xs[r][kk];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(zs[r][jq + 8 * q], x, acc[q]);
        }
        if (db && blockIdx.x == 0 && tid < kTN)
            for (int r = 0; r < kTK; ++r) bsum += zs[r][tid];
        __syncthreads();
    }
    if (k0 + kk < k)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = j0 + jq + 8 * q;
            if (j < n) atomicAdd(dw + (int64_t)j * ld_dw + k0 + kk, acc[q]);
        }
    if (db && blockIdx.x == 0 && tid < kTN && j0 + tid < n) atomicAdd(db + j0 + tid, bsum);
}

// dx[b, k] = sum_j dz[b, j] W[j, k].  pair mode (ia != NULL): scattered into d_emb[ia[b], k] / d_emb[ib[b], k - half]
// with atomics.  Otherwise written to dx and, when bn_stats != NULL, the two column sums of the BatchNorm backward are
// accumulated: bn_stats[k] += dx, bn_stats[kdim + k] += dx * xhat with xhat = (a[b, k] - mean[k]) * rstd[k].
__global__ void __launch_bounds__(kFcThreads) fc_bwd_input_kernel(const float* __restrict__ dz, int64_t ld_dz, int64_t m,
                                                                  int n, const float* __restrict__ w, int64_t ldw, int k,
                                                                  float* __restrict__ dx, int64_t ld_dx,
                                                                  const int64_t* __restrict__ ia,
                                                                  const int64_t* __restrict__ ib, int half,
                                                                  const float* __restrict__ a, int64_t ld_a,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  double* __restrict__ bn_stats) {
    __shared__ float zs[kTM][kTK + 1];     // [b][j chunk]
    __shared__ float ws[kTK][kTN + 1];     // [j chunk][k]
    __shared__ float cs[2][kTN];
    const int tid = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * kTM;
    const int k0 = blockIdx.y * kTN;
    const int r0 = tid >> 4, c0 = tid & 15;
    float acc[2][4] = {};
    for (int jj = 0; jj < n; jj += kTK) {
        for (int i = tid; i < kTM * kTK; i += kFcThreads) {
            const int r = i / kTK, c = i % kTK;
            zs[r][c] = (b0 + r < m && jj + c < n) ? __ldg(dz + (b0 + r) * ld_dz + jj + c) : 0.f;
        }
        for (int i = tid; i < kTK * kTN; i += kFcThreads) {
            const int r = i / kTN, c = i % kTN;
            ws[r][c] = (jj + r < n && k0 + c < k) ? __ldg(w + (int64_t)(jj + r) * ldw + k0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < kTK; ++c) {
            const float z0 = zs[r0][c], z1 = zs[r0 + 16][c];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float wv = ws[c][c0 + 16 * q];
                acc[0][q] = fmaf(z0, wv, acc[0][q]);
                acc[1][q] = fmaf(z1, wv, acc[1][q]);
            }
        }
        __syncthreads();
    }
    if (tid < 2 * kTN) cs[tid / kTN][tid % kTN] = 0.f;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int kc = k0 + c0 + 16 * q;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int64_t b = b0 + r0 + 16 * rr;
            if (b >= m || kc >= k) continue;
            const float v = acc[rr][q];
            if (ia) {
                const int64_t row = kc < half ? ia[b] : ib[b];
                atomicAdd(dx + row * ld_dx + (kc < half ? kc : kc - half), v);
            } else {
                dx[b * ld_dx + kc] = v;
                if (bn_stats) {
                    const float xhat = (__ldg(a + b * ld_a + kc) - __ldg(mean + kc)) * __ldg(rstd + kc);
                    s1 += v;
                    s2 = fmaf(v, xhat, s2);
                }
            }
        }
        if (bn_stats && kc < k) {
            atomicAdd(&cs[0][c0 + 16 * q], s1);
            atomicAdd(&cs[1][c0 + 16 * q], s2);
        }
    }
    if (bn_stats) {
        __syncthreads();
        if (tid < kTN && k0 + tid < k) {
            atomicAdd(bn_stats + k0 + tid, (double)cs[0][tid]);
            atomicAdd(bn_stats + k + k0 + tid, (double)cs[1][tid]);
        }
    }
}

// BatchNorm (training statistics) + ReLU backward, elementwise given the two column sums s1 = sum dy, s2 = sum dy xhat:
//   d a = gamma rstd (dy - s1 / m - xhat s2 / m) masked by the ReLU that produced a (a > 0);  dgamma = s2, dbeta = s1.
// eval statistics (training == 0): d a = gamma rstd dy.
__global__ void bn_relu_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ a, int64_t ld_a,
                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                   const float* __restrict__ gamma, const double* __restrict__ stats, int64_t m, int k,
                                   int training, float* __restrict__ dz, int64_t ld_dz, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta) {
    const int64_t total = m * k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / k;
        const int c = (int)(i - b * k);
        const float av = a[b * ld_a + c];
        const float xhat = (av - mean[c]) * rstd[c];
        float g = dy[b * ld_dy + c];
        if (training) g = g - (float)(stats[c] / (double)m) - xhat * (float)(stats[k + c] / (double)m);
        dz[b * ld_dz + c] = av > 0.f ? gamma[c] * rstd[c] * g : 0.f;
        if (b == 0) {
            dgamma[c] += (float)stats[k + c];
            dbeta[c] += (float)stats[c];
        }
    }
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t m,
                                   float* __restrict__ dz) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m) dz[i] = dy[i] * y[i] * (1.f - y[i]);
}

// ---- TransE (model_bce.py:329-368): one warp per triple --------------------------------------------------------
__device__ __forceinline__ float neg_logsigmoid(float x) { return fmaxf(-x, 0.f) + log1pf(__expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(256) transe_loss_kernel(const float* __restrict__ emb, int64_t ld, int dim,
                                                          const float* __restrict__ rel, int64_t ld_rel,
                                                          const int64_t* __restrict__ h, const int64_t* __restrict__ r,
                                                          const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                                                          int64_t batch, float lambda, float* __restrict__ loss,
                                                          const float* __restrict__ grad_scale, float* __restrict__ d_emb,
                                                          int64_t ld_d, float* __restrict__ d_rel) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= batch) return;
    const float* hr = emb + h[i] * ld;
    const float* pr = emb + pos[i] * ld;
    const float* nr = emb + neg[i] * ld;
    const float* er = rel + r[i] * ld_rel;
    float ps = 0.f, ns = 0.f, q = 0.f;
    for (int c = lane; c < dim; c += 32) {
        const float a = __ldg(hr + c), e = __ldg(er + c), b = __ldg(pr + c), d = __ldg(nr + c);
        const float u = a + e - b, v = a + e - d;
        ps = fmaf(u, u, ps);
        ns = fmaf(v, v, ns);
        q += a * a + e * e + b * b + d * d;
    }
    ps = warp_sum(ps); ns = warp_sum(ns); q = warp_sum(q);
    const float inv_b = 1.f / (float)batch;
    const float x = ns - ps;
    if (loss && lane == 0) atomicAdd(loss, (neg_logsigmoid(x) + lambda * 0.5f * q) * inv_b);
    if (d_emb) {
        const float up = grad_scale ? __ldg(grad_scale) : 1.f;
        const float s = -sigmoid_f(-x) * inv_b * up;        // d loss / d (neg_score - pos_score)
        const float l = lambda * inv_b * up;
        float* dh = d_emb + h[i] * ld_d;
        float* dp = d_emb + pos[i] * ld_d;
        float* dn = d_emb + neg[i] * ld_d;
        float* de = d_rel + r[i] * (int64_t)dim;
        for (int c = lane; c < dim; c += 32) {
            const float a = __ldg(hr + c), e = __ldg(er + c), b = __ldg(pr + c), d = __ldg(nr + c);
            const float u = a + e - b, v = a + e - d;
            const float common = 2.f * s * (v - u);
            atomicAdd(dh + c, common + l * a);
            atomicAdd(de + c, common + l * e);
            atomicAdd(dp + c, 2.f * s * u + l * b);
            atomicAdd(dn + c, -2.f * s * v + l * d);
        }
    }
}

inline FcInput make_input(const float* src, int64_t ld, const int64_t* ia, const int64_t* ib, int half,
                          const float* scale, const float* shift) {
    return FcInput{src, ld, ia, ib, half, scale, shift};
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_mlp_fc_fwd(const float* in, int64_t ld_in, const int64_t* rows_a, const int64_t* rows_b, int32_t half,
                              const float* in_scale, const float* in_shift, int64_t m, int32_t k, const float* w,
                              int64_t ldw, const float* bias, int32_t n, int32_t act, float* out, int64_t ld_out,
                              double* stats, void* stream_) {
    LKG_REQUIRE(in && w && out && m >= 0 && k > 0 && n > 0 && ld_out >= n && ldw >= k && act >= 0 && act <= 2,
                "bad fc arguments");
    LKG_REQUIRE((rows_a == nullptr) == (rows_b == nullptr) && (!rows_a || (half > 0 && half < k && !in_scale)),
                "pair mode needs both row lists and 0 < half < k");
    LKG_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "affine mode needs scale and shift");
    if (m == 0) return LKG_OK;
    dim3 grid((unsigned)((m + kTM - 1) / kTM), (unsigned)((n + kTN - 1) / kTN));
    fc_fwd_kernel<<<grid, kFcThreads, 0, (cudaStream_t)stream_>>>(make_input(in, ld_in, rows_a, rows_b, half, in_scale, in_shift),
                                                                  m, k, w, ldw, bias, n, act, out, ld_out, stats);
    LKG_LAUNCH_CHECK("fc_fwd_kernel");
    return LKG_OK;
}

extern "C" int lkg_bn_finalize(const double* stats, int64_t m, int32_t n, const float* gamma, const float* beta, float eps,
                               float momentum, float* running_mean, float* running_var, int32_t training, float* scale,
                               float* shift, float* mean_out, float* rstd_out, void* stream_) {
    LKG_REQUIRE(gamma && beta && scale && shift && n > 0 && eps > 0.f, "bad BatchNorm arguments");
    LKG_REQUIRE(training ? (stats != nullptr && m > 0) : (running_mean && running_var),
                "training needs batch statistics, evaluation the running buffers");
    bn_finalize_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_>>>(stats, m, n, gamma, beta, eps, momentum,
                                                                           running_mean, running_var, training, scale,
                                                                           shift, mean_out, rstd_out);
    LKG_LAUNCH_CHECK("bn_finalize_kernel");
    return LKG_OK;
}

extern "C" int lkg_mlp_fc_bwd_weight(const float* dz, int64_t ld_dz, const float* in, int64_t ld_in, const int64_t* rows_a,
                                     const int64_t* rows_b, int32_t half, const float* in_scale, const float* in_shift,
                                     int64_t m, int32_t k, int32_t n, float* dw, int64_t ld_dw, float* db, void* stream_) {
    LKG_REQUIRE(dz && in && dw && m >= 0 && k > 0 && n > 0 && ld_dw >= k && ld_dz >= n, "bad fc backward arguments");
    LKG_REQUIRE((rows_a == nullptr) == (rows_b == nullptr) && (in_scale == nullptr) == (in_shift == nullptr),
                "bad input transform");
    if (m == 0) return LKG_OK;
    int splits = (int)((m + 255) / 256);
    if (splits > 64) splits = 64;
    dim3 grid((unsigned)((k + kTM - 1) / kTM), (unsigned)((n + kTN - 1) / kTN), (unsigned)splits);
    fc_bwd_weight_kernel<<<grid, kFcThreads, 0, (cudaStream_t)stream_>>>(
        dz, ld_dz, make_input(in, ld_in, rows_a, rows_b, half, in_scale, in_shift), m, k, n, dw, ld_dw, db);
    LKG_LAUNCH_CHECK("fc_bwd_weight_kernel");
    return LKG_OK;
}

extern "C" int lkg_mlp_fc_bwd_input(const float* dz, int64_t ld_dz, int64_t m, int32_t n, const float* w, int64_t ldw,
                                    int32_t k, float* dx, int64_t ld_dx, const int64_t* rows_a, const int64_t* rows_b,
                                    int32_t half, const float* a, int64_t ld_a, const float* mean, const float* rstd,
                                    double* bn_stats, void* stream_) {
    LKG_REQUIRE(dz && w && dx && m >= 0 && n > 0 && k > 0 && ldw >= k && ld_dz >= n, "bad fc backward arguments");
    LKG_REQUIRE((rows_a == nullptr) == (rows_b == nullptr), "pair mode needs both row lists");
    LKG_REQUIRE(!bn_stats || (a && mean && rstd && !rows_a), "the BatchNorm sums need the saved activations");
    if (m == 0) return LKG_OK;
    dim3 grid((unsigned)((m + kTM - 1) / kTM), (unsigned)((k + kTN - 1) / kTN));
    fc_bwd_input_kernel<<<grid, kFcThreads, 0, (cudaStream_t)stream_>>>(dz, ld_dz, m, n, w, ldw, k, dx, ld_dx, rows_a,
                                                                        rows_b, half, a, ld_a, mean, rstd, bn_stats);
    LKG_LAUNCH_CHECK("fc_bwd_input_kernel");
    return LKG_OK;
}

extern "C" int lkg_bn_relu_bwd(const float* dy, int64_t ld_dy, const float* a, int64_t ld_a, const float* mean,
                               const float* rstd, const float* gamma, const double* stats, int64_t m, int32_t k,
                               int32_t training, float* dz, int64_t ld_dz, float* dgamma, float* dbeta, void* stream_) {
    LKG_REQUIRE(dy && a && mean && rstd && gamma && stats && dz && dgamma && dbeta && m > 0 && k > 0,
                "bad BatchNorm backward arguments");
    int64_t blocks = (m * k + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    bn_relu_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(dy, ld_dy, a, ld_a, mean, rstd, gamma, stats, m,
                                                                            k, training, dz, ld_dz, dgamma, dbeta);
    LKG_LAUNCH_CHECK("bn_relu_bwd_kernel");
    return LKG_OK;
}

extern "C" int lkg_sigmoid_bwd(const float* dy, const float* y, int64_t m, float* dz, void* stream_) {
    LKG_REQUIRE(dy && y && dz && m >= 0, "bad sigmoid backward arguments");
    if (m == 0) return LKG_OK;
    sigmoid_bwd_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(dy, y, m, dz);
    LKG_LAUNCH_CHECK("sigmoid_bwd_kernel");
    return LKG_OK;
}

extern "C" int lkg_transe_loss(const float* emb, int64_t ld_emb, int32_t dim, const float* relation, int64_t ld_rel,
                               const int64_t* h, const int64_t* r, const int64_t* pos, const int64_t* neg, int64_t batch,
                               float l2_lambda, float* loss, const float* grad_scale, float* d_emb, int64_t ld_d,
                               float* d_relation, void* stream_) {
    LKG_REQUIRE(emb && relation && h && r && pos && neg && (loss || d_emb) && dim > 0 && batch > 0, "bad TransE arguments");
    LKG_REQUIRE(!d_emb || d_relation, "the backward needs both gradient buffers");
    const int64_t blocks = (batch * 32 + 255) / 256;
    transe_loss_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(emb, ld_emb, dim, relation, ld_rel, h, r, pos,
                                                                            neg, batch, l2_lambda, loss, grad_scale, d_emb,
                                                                            ld_d, d_relation);
    LKG_LAUNCH_CHECK("transe_loss_kernel");
    return LKG_OK;
}
