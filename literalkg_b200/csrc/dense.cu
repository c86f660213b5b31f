// Dense products of the path:  C[M,N] = epilogue(A[M,K] @ B[N,K]^T), fp32 accumulate.
//   lkg_linear_fwd  : act(A B^T + bias)                       linear_gat (model.py:309-310), residual pre-projection
//   lkg_gate_fwd    : literal gate with tanh / sigmoid / mix   GateMul / Gate (gate.py:22-28, 45-51)
//   lkg_score       : emb[heads] @ emb[tails]^T + global min/max (model.py:473-486, 490)
// A is a K-concatenation of up to LKG_MAX_SEGMENTS row-major sources (the virtual torch.cat of the
// reference) with optional row gather; B is in torch Linear layout [out, in] so no transposes are
// materialised.  This file is the exact-fp32 CUDA-core engine (128x64x16 tiles, 8x4 register tile).
#include "common.cuh"

namespace lkg {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, THREADS = 256;
constexpr int AS_LD = BM + 4, BS_LD = BN + 4;

struct Operand {
    lkg_operand a;
};

struct LinearEpi {
    const float* bias;
    int act;
    float* out;
    int64_t ldo;
    __device__ void operator()(int64_t m, int n0, const float (&acc)[TN], int n_total) const {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + j;
            if (n < n_total) {
                float v = acc[j] + (bias ? __ldg(bias + n) : 0.f);
                if (act == LKG_ACT_LEAKY_RELU) v = leaky(v);
                out[m * ldo + n] = v;
            }
        }
    }
};

struct GateEpi {
    const float* bias_pair;
    const float* x_ent;
    int64_t ld_ent;
    float* out;
    int64_t ldo;
    __device__ void operator()(int64_t m, int n0, const float (&acc)[TN], int n_total) const {
#pragma unroll
        for (int j = 0; j < TN; j += 2) {
            const int n = n0 + j;   // even: g pre-activation, n + 1: gate pre-activation
            if (n + 1 < n_total) {
                const int c = n >> 1;
                const float g = tanh_acc(acc[j] + __ldg(bias_pair + n));
                const float z = sigmoid_acc(acc[j + 1] + __ldg(bias_pair + n + 1));
                const float e = __ldg(x_ent + m * ld_ent + c);
                out[m * ldo + c] = (1.f - z) * e + z * g;
            }
        }
    }
};

__device__ __forceinline__ uint32_t order_enc(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct ScoreEpi {
    float* out;
    int64_t ldo;
    uint32_t* minmax;   // ordered-encoded {min, max}, nullable
    __device__ void operator()(int64_t m, int n0, const float (&acc)[TN], int n_total) const {
#pragma unroll
        for (int j = 0; j < TN; ++j)
            if (n0 + j < n_total) out[m * ldo + n0 + j] = acc[j];
    }
};

template <class Epi, bool kMinMax>
__global__ void __launch_bounds__(THREADS)
gemm_tn_kernel(lkg_operand a, int64_t m_total, const float* __restrict__ b, int64_t ldb,
               const int64_t* __restrict__ b_rows, int n_total, Epi epi, uint32_t* minmax) {
    __shared__ __align__(16) float As[BK][AS_LD];
    __shared__ __align__(16) float Bs[BK][BS_LD];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);   // 0..15 -> columns
    const int ty = tid / (BN / TN);   // 0..15 -> rows
    const int64_t m0 = (int64_t)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // loader mapping: A tile 128 x 16 -> thread loads 8 consecutive k of one row;
    //                 B tile  64 x 16 -> thread loads 4 consecutive k of one row
    const int a_row = tid >> 1, a_k = (tid & 1) * 8;
    const int b_row = tid >> 2, b_k = (tid & 3) * 4;
    const int64_t am = m0 + a_row;
    const int64_t a_src_row = am < m_total ? (a.rows ? a.rows[am] : am) : -1;
    const int bn = n0 + b_row;
    const int64_t b_src_row = bn < n_total ? (b_rows ? b_rows[bn] : bn) : -1;

    int k_base = 0;
    for (int seg = 0; seg < a.n_segments; ++seg) {
        const float* ap = a.ptr[seg];
        const int64_t lda = a.ld[seg];
        const int kseg = a.k[seg];
        for (int k0 = 0; k0 < kseg; k0 += BK) {
            float av[8], bv[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = k0 + a_k + i;
                av[i] = (a_src_row >= 0 && k < kseg) ? __ldg(ap + a_src_row * lda + k) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + b_k + i;
                bv[i] = (b_src_row >= 0 && k < kseg) ? __ldg(b + b_src_row * ldb + k_base + k) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) As[a_k + i][a_row] = av[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[b_k + i][b_row] = bv[i];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
                const float ar[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float br[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
            }
        }
        k_base += kseg;
    }

    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ty * TM + i;
        if (m < m_total) {
            epi(m, n0 + tx * TN, acc[i], n_total);
            if (kMinMax) {
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (n0 + tx * TN + j < n_total) {
                        lo = fminf(lo, acc[i][j]);
                        hi = fmaxf(hi, acc[i][j]);
                    }
            }
        }
    }
    if (kMinMax && minmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(kFull, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(kFull, hi, o));
        }
        if ((tid & 31) == 0) {
            if (lo <= hi) {
                atomicMin(minmax, order_enc(lo));
                atomicMax(minmax + 1, order_enc(hi));
            }
        }
    }
}

int check_operand(const lkg_operand* a) {
    LKG_REQUIRE(a != nullptr, "operand is null");
    LKG_REQUIRE(a->n_segments >= 1 && a->n_segments <= LKG_MAX_SEGMENTS, "bad segment count %d", a->n_segments);
    for (int i = 0; i < a->n_segments; ++i) {
        LKG_REQUIRE(a->ptr[i] != nullptr && a->k[i] > 0 && a->ld[i] >= a->k[i], "bad operand segment %d", i);
    }
    return LKG_OK;
}

__global__ void minmax_reset_kernel(uint32_t* mm) {
    mm[0] = 0xffffffffu;
    mm[1] = 0u;
}

__global__ void threshold_kernel(const float* __restrict__ s, int64_t lds, int64_t rows, int64_t cols,
                                 const uint32_t* __restrict__ mm, float milestone, int32_t* __restrict__ pred,
                                 int64_t ldp) {
    auto dec = [](uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); };
    const float lo = dec(mm[0]), hi = dec(mm[1]);
    const float range = hi - lo;
    const int64_t total = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const float v = (s[r * lds + c] - lo) / range;   // same op order as model.py:490
        pred[r * ldp + c] = v > milestone ? 1 : 0;
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_linear_fwd(const lkg_operand* a, int64_t m, const float* b, int64_t ldb, int32_t n,
                              const float* bias, int32_t activation, float* out, int64_t ldo, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_operand(a)) return rc;
    LKG_REQUIRE(b && out && m >= 0 && n > 0 && ldo >= n, "bad linear arguments");
    if (m == 0) return LKG_OK;
    LinearEpi epi{bias, activation, out, ldo};
    dim3 grid((n + BN - 1) / BN, (unsigned)((m + BM - 1) / BM));
    gemm_tn_kernel<LinearEpi, false><<<grid, THREADS, 0, stream>>>(*a, m, b, ldb, nullptr, n, epi, nullptr);
    LKG_LAUNCH_CHECK("gemm_tn_kernel<linear>");
    return LKG_OK;
}

extern "C" int lkg_gate_fwd(const lkg_operand* x, int64_t m, const float* w_pair, int64_t ldw,
                            const float* bias_pair, int32_t dim, const float* x_ent, int64_t ld_ent,
                            float* out, int64_t ldo, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_operand(x)) return rc;
    LKG_REQUIRE(w_pair && bias_pair && x_ent && out && dim > 0 && ldo >= dim, "bad gate arguments");
    if (m == 0) return LKG_OK;
    GateEpi epi{bias_pair, x_ent, ld_ent, out, ldo};
    const int n = 2 * dim;
    dim3 grid((n + BN - 1) / BN, (unsigned)((m + BM - 1) / BM));
    gemm_tn_kernel<GateEpi, false><<<grid, THREADS, 0, stream>>>(*x, m, w_pair, ldw, nullptr, n, epi, nullptr);
    LKG_LAUNCH_CHECK("gemm_tn_kernel<gate>");
    return LKG_OK;
}

extern "C" int lkg_minmax_reset(uint32_t* minmax_dev, void* stream_) {
    LKG_REQUIRE(minmax_dev != nullptr, "minmax is null");
    minmax_reset_kernel<<<1, 1, 0, (cudaStream_t)stream_>>>(minmax_dev);
    LKG_LAUNCH_CHECK("minmax_reset_kernel");
    return LKG_OK;
}

extern "C" int lkg_score(const float* emb, int64_t ld_emb, int32_t dim, const int64_t* heads, int64_t n_heads,
                         const int64_t* tails, int64_t n_tails, float* scores, int64_t ld_scores,
                         uint32_t* minmax_dev, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(emb && scores && dim > 0 && ld_emb >= dim, "bad score arguments");
    LKG_REQUIRE(n_heads >= 0 && n_tails >= 0 && n_tails < (1ll << 31) && ld_scores >= n_tails, "bad score shape");
    if (n_heads == 0 || n_tails == 0) return LKG_OK;
    lkg_operand a{};
    a.n_segments = 1;
    a.ptr[0] = emb;
    a.ld[0] = ld_emb;
    a.k[0] = dim;
    a.rows = heads;
    ScoreEpi epi{scores, ld_scores, minmax_dev};
    dim3 grid((unsigned)((n_tails + BN - 1) / BN), (unsigned)((n_heads + BM - 1) / BM));
    if (minmax_dev)
        gemm_tn_kernel<ScoreEpi, true><<<grid, THREADS, 0, stream>>>(a, n_heads, emb, ld_emb, tails, (int)n_tails, epi, minmax_dev);
    else
        gemm_tn_kernel<ScoreEpi, false><<<grid, THREADS, 0, stream>>>(a, n_heads, emb, ld_emb, tails, (int)n_tails, epi, nullptr);
    LKG_LAUNCH_CHECK("gemm_tn_kernel<score>");
    return LKG_OK;
}

extern "C" int lkg_predict_threshold(const float* scores, int64_t ld_scores, int64_t n_heads, int64_t n_tails,
                                     const uint32_t* minmax_dev, float milestone, int32_t* pred, int64_t ld_pred,
                                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(scores && minmax_dev && pred, "null argument");
    const int64_t total = n_heads * n_tails;
    if (total == 0) return LKG_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    threshold_kernel<<<(int)blocks, 256, 0, stream>>>(scores, ld_scores, n_heads, n_tails, minmax_dev, milestone, pred, ld_pred);
    LKG_LAUNCH_CHECK("threshold_kernel");
    return LKG_OK;
}
