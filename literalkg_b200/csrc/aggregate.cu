// One aggregator layer, forward: CSR SpMM fused with the combine, LeakyReLU, LayerNorm, dropout mask
// and the L2-normalised copy for the concat buffer.
//
// Replaces Aggregator.forward + residual_connection (model.py:90-164) and the F.normalize of
// model.py:305.  The reference runs cuSPARSE SpMM, then 2-4 dense N x d x d GEMMs, ~10 elementwise
// kernels and a LayerNorm, each a full pass over N x d.  Here one warp owns kRows head rows:
//   phase 1  side = sum_j A[row, j] * ego[col_j]   128-bit streaming gathers, kUnroll neighbours in flight
//   phase 2  u-vectors (side | ego+side | ego | ego*side) staged in shared memory, then the folded
//            combine  o = u @ P + r[row]  with lane = output channel (P lives in shared memory)
//   phase 3  leaky / add / LayerNorm / mask / L2 normalise, all in registers + warp shuffles
// Folding (host side): linear(res(hi)) = hi @ P + h0 @ Q + c, P = (1-a) M W^T.
// HBM bound: algorithmic bytes = nnz*(4 col + 4 val + 4 d_in) + N*(4 d_in + 8 + r terms + 2*4*d_out).
#include <cuda_fp16.h>

#include "common.cuh"

namespace lkg {
namespace {

constexpr int kRows = 2;  // head rows per warp per step (P is read from shared memory once per kRows rows)

enum Mode { kOneTerm = 0, kTwoTerms = 1, kBi = 2 };

struct AggParams {
    lkg_graph g;
    const float* a_val;
    const float* ego;
    int64_t ld_ego;
    int d_in, d_out, nvec;
    const float* p0;   // term 0 matrix [d_in, d_out]
    const float* p1;   // term 1 matrix (kTwoTerms: side matrix; kBi: product matrix)
    int sum_ego;       // term 0 vector: 1 -> ego + side, 0 -> side   (kOneTerm / kBi);  kTwoTerms: term0 = ego, term1 = side
    const float* r1;
    const float* r2;
    int64_t ld_r;
    const float* ln_w;
    const float* ln_b;
    const float* mask;
    float* x_out;
    int64_t ld_x;
    float* xn_out;
    int64_t ld_xn;
    __half* xn_planes;          // optional scaled hi/lo fp16 copy of xn (operand of the linear_gat tensor-core GEMM)
    int64_t ld_planes, plane_stride;
    const float* xn_rec;        // its scale record (|xn| <= 1)
    int* counter;
    int p_stride;      // padded row stride (floats) of P in shared memory
    int64_t local_row_base;   // r1 / r2 / mask / xn_out / xn_planes are indexed by (row - local_row_base): the
                              // row partition's local buffers; ego and x_out are indexed by the global row
};

template <int S, int NC, int MODE>
__global__ void __launch_bounds__(512, 1) aggregate_kernel(AggParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int UNROLL = S == 1 ? 8 : 4;   // narrow rows: more neighbours in flight
    constexpr int NT = MODE == kOneTerm ? 1 : 2;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int d_in = p.d_in, d_out = p.d_out, nvec = p.nvec;
    const int ps = p.p_stride;
    float* sp0 = smem;
    float* sp1 = smem + (size_t)d_in * ps;
    float* stage = smem + (size_t)NT * d_in * ps + (size_t)warp * (kRows * NT * d_in);

    // stage the folded matrices once per CTA
    for (int i = threadIdx.x; i < d_in * d_out; i += blockDim.x) {
        const int d = i / d_out, c = i - d * d_out;
        sp0[d * ps + c] = p.p0[i];
        if (NT == 2) sp1[d * ps + c] = p.p1[i];
    }
    __syncthreads();
    (void)nwarps;

    const int n = (int)(p.g.row_end - p.g.row_begin);      // rows of the partition, taken in row_order
    const int row0 = (int)p.g.row_begin;
    auto row_of = [&](int i) { return p.g.row_order ? __ldg(p.g.row_order + i) : row0 + i; };
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(p.counter, kRows);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;

        // ---- phase 1: SpMM + u-vector staging ------------------------------------------------
#pragma unroll
        for (int rr = 0; rr < kRows; ++rr) {
            if (base + rr >= n) break;
            const int row = row_of(base + rr);
            float4 side[S];
#pragma unroll
            for (int s = 0; s < S; ++s) side[s] = make_float4(0, 0, 0, 0);
            const int u0 = p.g.rowptr[row], u1 = p.g.rowptr[row + 1];
            for (int u = u0; u < u1; u += UNROLL) {
                int cl[UNROLL];
                float av[UNROLL];
#pragma unroll
                for (int j = 0; j < UNROLL; ++j) {
                    const bool live = u + j < u1;
                    cl[j] = live ? __ldg(p.g.col + u + j) : -1;
                    av[j] = live ? __ldg(p.a_val + u + j) : 0.f;
                }
                float4 x[UNROLL][S];
#pragma unroll
                for (int j = 0; j < UNROLL; ++j) {
                    const float* src = p.ego + (int64_t)(cl[j] < 0 ? 0 : cl[j]) * p.ld_ego;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const int v = lane + 32 * s;
                        x[j][s] = (cl[j] >= 0 && v < nvec) ? ldg_stream4(src + 4 * v) : make_float4(0, 0, 0, 0);
                    }
                }
#pragma unroll
                for (int j = 0; j < UNROLL; ++j) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        side[s].x = fmaf(av[j], x[j][s].x, side[s].x);
                        side[s].y = fmaf(av[j], x[j][s].y, side[s].y);
                        side[s].z = fmaf(av[j], x[j][s].z, side[s].z);
                        side[s].w = fmaf(av[j], x[j][s].w, side[s].w);
                    }
                }
            }
            const float* erow = p.ego + (int64_t)row * p.ld_ego;
            float* st = stage + rr * NT * d_in;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int v = lane + 32 * s;
                if (v < nvec) {
                    const float4 sd = side[s];
                    float4 eg = make_float4(0, 0, 0, 0);
                    if (MODE != kOneTerm || p.sum_ego) eg = __ldg(reinterpret_cast<const float4*>(erow) + v);
                    float4 t0, t1;
                    if (MODE == kTwoTerms) {
                        t0 = eg;
                        t1 = sd;
                    } else {
                        t0 = p.sum_ego ? make_float4(eg.x + sd.x, eg.y + sd.y, eg.z + sd.z, eg.w + sd.w) : sd;
                        t1 = make_float4(eg.x * sd.x, eg.y * sd.y, eg.z * sd.z, eg.w * sd.w);
                    }
                    reinterpret_cast<float4*>(st)[v] = t0;
                    if (NT == 2) reinterpret_cast<float4*>(st + d_in)[v] = t1;
                }
            }
        }
        __syncwarp();

        // ---- phase 2: folded combine, lane = output channel ----------------------------------
        float acc1[kRows][NC], acc2[kRows][NC];
#pragma unroll
        for (int rr = 0; rr < kRows; ++rr) {
            const int64_t lrow = row_of(min(base + rr, n - 1)) - p.local_row_base;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int ch = lane + 32 * c;
                acc1[rr][c] = (ch < d_out && p.r1) ? __ldg(p.r1 + lrow * p.ld_r + ch) : 0.f;
                acc2[rr][c] = (MODE == kBi && ch < d_out && p.r2) ? __ldg(p.r2 + lrow * p.ld_r + ch) : 0.f;
            }
        }
        for (int d4 = 0; d4 < nvec; ++d4) {
            float4 u[kRows][NT];
#pragma unroll
            for (int rr = 0; rr < kRows; ++rr)
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    u[rr][t] = reinterpret_cast<const float4*>(stage + (rr * NT + t) * d_in)[d4];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int ch = lane + 32 * c;
                const int chc = ch < d_out ? ch : 0;
                float w0[4], w1[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w0[i] = sp0[(4 * d4 + i) * ps + chc];
                    w1[i] = NT == 2 ? sp1[(4 * d4 + i) * ps + chc] : 0.f;
                }
#pragma unroll
                for (int rr = 0; rr < kRows; ++rr) {
                    float a = acc1[rr][c];
                    a = fmaf(u[rr][0].x, w0[0], a);
                    a = fmaf(u[rr][0].y, w0[1], a);
                    a = fmaf(u[rr][0].z, w0[2], a);
                    a = fmaf(u[rr][0].w, w0[3], a);
                    if (MODE == kTwoTerms) {
                        a = fmaf(u[rr][1].x, w1[0], a);
                        a = fmaf(u[rr][1].y, w1[1], a);
                        a = fmaf(u[rr][1].z, w1[2], a);
                        a = fmaf(u[rr][1].w, w1[3], a);
                    }
                    acc1[rr][c] = a;
                    if (MODE == kBi) {
                        float b = acc2[rr][c];
                        b = fmaf(u[rr][1].x, w1[0], b);
                        b = fmaf(u[rr][1].y, w1[1], b);
                        b = fmaf(u[rr][1].z, w1[2], b);
                        b = fmaf(u[rr][1].w, w1[3], b);
                        acc2[rr][c] = b;
                    }
                }
            }
        }
        __syncwarp();   // staging buffer is reused by the next step

        // ---- phase 3: activation, LayerNorm, mask, L2 normalise --------------------------------
#pragma unroll
        for (int rr = 0; rr < kRows; ++rr) {
            if (base + rr >= n) break;
            const int row = row_of(base + rr);
            const int64_t lrow = row - p.local_row_base;
            float emb[NC];
            float s1 = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int ch = lane + 32 * c;
                float e = leaky(acc1[rr][c]);
                if (MODE == kBi) e += leaky(acc2[rr][c]);
                emb[c] = ch < d_out ? e : 0.f;
                s1 += emb[c];
            }
            const float mean = warp_sum(s1) / (float)d_out;
            float s2 = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int ch = lane + 32 * c;
                const float dlt = ch < d_out ? emb[c] - mean : 0.f;
                s2 = fmaf(dlt, dlt, s2);
            }
            const float rstd = rsqrtf(warp_sum(s2) / (float)d_out + 1e-5f);
            float sq = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int ch = lane + 32 * c;
                float x = 0.f;
                if (ch < d_out) {
                    x = (emb[c] - mean) * rstd * __ldg(p.ln_w + ch) + __ldg(p.ln_b + ch);
                    if (p.mask) x *= __ldg(p.mask + lrow * d_out + ch);
                    p.x_out[(int64_t)row * p.ld_x + ch] = x;
                }
                emb[c] = x;
                sq = fmaf(x, x, sq);
            }
            if (p.xn_out || p.xn_planes) {
                const float inv = 1.f / fmaxf(sqrtf(warp_sum(sq)), 1e-12f);
                const float pscale = p.xn_planes ? __ldg(p.xn_rec + 1) : 1.f;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int ch = lane + 32 * c;
                    if (ch < d_out) {
                        const float xn = emb[c] * inv;
                        if (p.xn_out) p.xn_out[lrow * p.ld_xn + ch] = xn;
                        if (p.xn_planes) {
                            const float xs = xn * pscale;
                            const __half h = __float2half_rn(xs);
                            const __half l = __float2half_rn(xs - __half2float(h));
                            p.xn_planes[lrow * p.ld_planes + ch] = h;
                            p.xn_planes[p.plane_stride + lrow * p.ld_planes + ch] = l;
                        }
                    }
                }
            }
        }
    }
}

// ---- narrow rows (d_in = 16 / 32 / 64): LPR = d_in / 4 lanes own one head row, 32 / LPR rows per warp ----
// Layers >= 2 gather 128-byte rows: with one warp per row 24 of the 32 lanes idle and the kernel is latency
// bound (ncu r01a: 0.9 TB/s, 22 % occupancy).  Here every lane holds one float4 of its row; the neighbour list is
// loaded LPR entries at a time (one coalesced load per group) and broadcast with width-LPR shuffles, all LPR
// gathers of a chunk in flight together.  Batches come from the degree-sorted row_order, so the rows that share
// a warp have similar lengths; a batch whose heaviest row exceeds kCoopDegree is processed row by row with the
// whole warp splitting the row's neighbour list (then summed across groups).  The folded combine keeps the
// u-vectors in registers: lane gl owns the output channels [4 gl, 4 gl + 4) (+ 4 LPR i), u[d] arrives by shuffle
// from the lane that holds it and P rows are read from shared memory as float4.
constexpr int kCoopDegree = 96;

__device__ __forceinline__ float4 f4_fma(float a, float4 w, float4 c) {
    return make_float4(fmaf(a, w.x, c.x), fmaf(a, w.y, c.y), fmaf(a, w.z, c.z), fmaf(a, w.w, c.w));
}

template <int LPR, int MODE, int CPL>   // CPL = float4 channel chunks per lane = ceil(d_out / (4 LPR))
__global__ void __launch_bounds__(256) aggregate_narrow_kernel(AggParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int RPW = 32 / LPR;
    constexpr int NT = MODE == kOneTerm ? 1 : 2;
    constexpr int U = LPR < 8 ? LPR : 8;
    const int lane = threadIdx.x & 31;
    const int grp = lane / LPR, gl = lane % LPR;
    const int d_in = 4 * LPR, d_out = p.d_out;
    float* sp0 = smem;
    float* sp1 = smem + d_in * d_out;
    for (int i = threadIdx.x; i < d_in * d_out; i += blockDim.x) {
        sp0[i] = p.p0[i];
        if (NT == 2) sp1[i] = p.p1[i];
    }
    __syncthreads();

    const int n = (int)(p.g.row_end - p.g.row_begin);
    const int row0 = (int)p.g.row_begin;
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(p.counter, RPW);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        const bool live = base + grp < n;
        const int row = live ? (p.g.row_order ? __ldg(p.g.row_order + base + grp) : row0 + base + grp) : 0;
        const int64_t lrow = row - p.local_row_base;
        const int u0 = live ? __ldg(p.g.rowptr + row) : 0;
        const int u1 = live ? __ldg(p.g.rowptr + row + 1) : 0;
        int max_deg = u1 - u0;
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) max_deg = max(max_deg, __shfl_xor_sync(kFull, max_deg, o));
        max_deg = __shfl_sync(kFull, max_deg, 0);
        const bool coop = RPW > 1 && max_deg > kCoopDegree;

        // ---- phase 1: side = sum_j A[row, j] * ego[col_j], this lane's float4 -------------------------
        float4 side = make_float4(0, 0, 0, 0);
        const float* ego_l = p.ego + 4 * gl;
        for (int pass = 0; pass < (coop ? RPW : 1); ++pass) {
            // group mode: every group walks its own row, LPR neighbours per step; coop mode: the warp walks the
            // row of group `pass`, 32 neighbours per step (group g takes neighbours [g LPR, (g+1) LPR) of the step)
            const int b0 = coop ? __shfl_sync(kFull, u0, pass * LPR) : u0;
            const int b1 = coop ? __shfl_sync(kFull, u1, pass * LPR) : u1;
            const int my = coop ? lane : gl;
            const int step = coop ? 32 : LPR;
            const int trips = coop ? (b1 - b0 + 31) / 32 : (max_deg + LPR - 1) / LPR;
            float4 acc = make_float4(0, 0, 0, 0);
            for (int it = 0; it < trips; ++it) {
                const int u = b0 + it * step + my;
                const bool ok = u < b1;
                const int cl = ok ? __ldg(p.g.col + u) : -1;
                const float av = ok ? __ldg(p.a_val + u) : 0.f;
#pragma unroll
                for (int j0 = 0; j0 < LPR; j0 += U) {
                    float4 x[U];
                    float a[U];
#pragma unroll
                    for (int j = 0; j < U; ++j) {
                        const int c = __shfl_sync(kFull, cl, j0 + j, LPR);
                        a[j] = __shfl_sync(kFull, av, j0 + j, LPR);
                        x[j] = c >= 0 ? ldg_stream4(ego_l + (int64_t)c * p.ld_ego) : make_float4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int j = 0; j < U; ++j) acc = f4_fma(a[j], x[j], acc);
                }
            }
            if (coop) {
#pragma unroll
                for (int o = LPR; o < 32; o <<= 1) {
                    acc.x += __shfl_xor_sync(kFull, acc.x, o);
                    acc.y += __shfl_xor_sync(kFull, acc.y, o);
                    acc.z += __shfl_xor_sync(kFull, acc.z, o);
                    acc.w += __shfl_xor_sync(kFull, acc.w, o);
                }
                if (grp == pass) side = acc;
            } else {
                side = acc;
            }
        }

        // ---- phase 2: folded combine in registers ------------------------------------------------------
        float4 eg = make_float4(0, 0, 0, 0);
        if (live && (MODE != kOneTerm || p.sum_ego))
            eg = __ldg(reinterpret_cast<const float4*>(p.ego + (int64_t)row * p.ld_ego) + gl);
        float4 t0, t1;
        if (MODE == kTwoTerms) {
            t0 = eg;
            t1 = side;
        } else {
            t0 = p.sum_ego ? make_float4(eg.x + side.x, eg.y + side.y, eg.z + side.z, eg.w + side.w) : side;
            t1 = make_float4(eg.x * side.x, eg.y * side.y, eg.z * side.z, eg.w * side.w);
        }
        float4 acc1[CPL], acc2[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int ch = 4 * (gl + LPR * i);
            const bool okc = live && ch < d_out;
            acc1[i] = (okc && p.r1) ? __ldg(reinterpret_cast<const float4*>(p.r1 + lrow * p.ld_r + ch))
                                    : make_float4(0, 0, 0, 0);
            acc2[i] = (MODE == kBi && okc && p.r2) ? __ldg(reinterpret_cast<const float4*>(p.r2 + lrow * p.ld_r + ch))
                                                   : make_float4(0, 0, 0, 0);
        }
#pragma unroll 2
        for (int sl = 0; sl < LPR; ++sl) {
            const float a0[4] = {__shfl_sync(kFull, t0.x, sl, LPR), __shfl_sync(kFull, t0.y, sl, LPR),
                                 __shfl_sync(kFull, t0.z, sl, LPR), __shfl_sync(kFull, t0.w, sl, LPR)};
            float a1[4] = {0.f, 0.f, 0.f, 0.f};
            if (NT == 2) {
                a1[0] = __shfl_sync(kFull, t1.x, sl, LPR);
                a1[1] = __shfl_sync(kFull, t1.y, sl, LPR);
                a1[2] = __shfl_sync(kFull, t1.z, sl, LPR);
                a1[3] = __shfl_sync(kFull, t1.w, sl, LPR);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int d = 4 * sl + c;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const int ch = 4 * (gl + LPR * i);
                    if (ch < d_out) {
                        const float4 w0 = *reinterpret_cast<const float4*>(sp0 + d * d_out + ch);
                        acc1[i] = f4_fma(a0[c], w0, acc1[i]);
                        if (NT == 2) {
                            const float4 w1 = *reinterpret_cast<const float4*>(sp1 + d * d_out + ch);
                            if (MODE == kTwoTerms) acc1[i] = f4_fma(a1[c], w1, acc1[i]);
                            else acc2[i] = f4_fma(a1[c], w1, acc2[i]);
                        }
                    }
                }
            }
        }

        // ---- phase 3: activation, LayerNorm, mask, L2 normalise (group-wide reductions) -----------------
        float4 emb[CPL];
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int ch = 4 * (gl + LPR * i);
            float4 e = make_float4(leaky(acc1[i].x), leaky(acc1[i].y), leaky(acc1[i].z), leaky(acc1[i].w));
            if (MODE == kBi) {
                e.x += leaky(acc2[i].x);
                e.y += leaky(acc2[i].y);
                e.z += leaky(acc2[i].z);
                e.w += leaky(acc2[i].w);
            }
            if (ch >= d_out) e = make_float4(0, 0, 0, 0);
            emb[i] = e;
            s1 += (e.x + e.y) + (e.z + e.w);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s1 += __shfl_xor_sync(kFull, s1, o);
        const float mean = s1 / (float)d_out;
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            if (4 * (gl + LPR * i) < d_out) {
                const float dx = emb[i].x - mean, dy = emb[i].y - mean, dz = emb[i].z - mean, dw = emb[i].w - mean;
                s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s2 += __shfl_xor_sync(kFull, s2, o);
        const float rstd = rsqrtf(s2 / (float)d_out + 1e-5f);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int ch = 4 * (gl + LPR * i);
            float4 x = make_float4(0, 0, 0, 0);
            if (ch < d_out) {
                const float4 lw = __ldg(reinterpret_cast<const float4*>(p.ln_w + ch));
                const float4 lb = __ldg(reinterpret_cast<const float4*>(p.ln_b + ch));
                x.x = (emb[i].x - mean) * rstd * lw.x + lb.x;
                x.y = (emb[i].y - mean) * rstd * lw.y + lb.y;
                x.z = (emb[i].z - mean) * rstd * lw.z + lb.z;
                x.w = (emb[i].w - mean) * rstd * lw.w + lb.w;
                if (p.mask && live) {
                    const float4 mk = __ldg(reinterpret_cast<const float4*>(p.mask + lrow * d_out + ch));
                    x.x *= mk.x; x.y *= mk.y; x.z *= mk.z; x.w *= mk.w;
                }
                if (live) *reinterpret_cast<float4*>(p.x_out + (int64_t)row * p.ld_x + ch) = x;
            }
            emb[i] = x;
            sq += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
        }
        if (p.xn_out || p.xn_planes) {
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(kFull, sq, o);
            const float inv = 1.f / fmaxf(sqrtf(sq), 1e-12f);
            const float pscale = p.xn_planes ? __ldg(p.xn_rec + 1) : 1.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int ch = 4 * (gl + LPR * i);
                if (ch < d_out && live) {
                    const float4 xn = make_float4(emb[i].x * inv, emb[i].y * inv, emb[i].z * inv, emb[i].w * inv);
                    if (p.xn_out) *reinterpret_cast<float4*>(p.xn_out + lrow * p.ld_xn + ch) = xn;
                    if (p.xn_planes) {
                        const float v[4] = {xn.x * pscale, xn.y * pscale, xn.z * pscale, xn.w * pscale};
                        __align__(8) __half h[4], l[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            h[c] = __float2half_rn(v[c]);
                            l[c] = __float2half_rn(v[c] - __half2float(h[c]));
                        }
                        __half* hp = p.xn_planes + lrow * p.ld_planes + ch;
                        *reinterpret_cast<uint2*>(hp) = *reinterpret_cast<const uint2*>(h);
                        *reinterpret_cast<uint2*>(hp + p.plane_stride) = *reinterpret_cast<const uint2*>(l);
                    }
                }
            }
        }
    }
}

template <int LPR, int MODE, int CPL>
int launch_narrow(const AggParams& p, cudaStream_t stream) {
    auto kern = aggregate_narrow_kernel<LPR, MODE, CPL>;
    const size_t smem = (size_t)(MODE == kOneTerm ? 1 : 2) * (4 * LPR) * p.d_out * sizeof(float);
    LKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    LKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) LKG_FAIL(LKG_ERR_UNSUPPORTED, "narrow aggregate kernel does not fit (smem %zu)", smem);
    kern<<<sm_count() * per_sm, 256, smem, stream>>>(p);
    LKG_LAUNCH_CHECK("aggregate_narrow_kernel");
    return LKG_OK;
}

template <int LPR, int CPL>
int dispatch_narrow(int mode, const AggParams& p, cudaStream_t stream) {
    switch (mode) {
        case kOneTerm: return launch_narrow<LPR, kOneTerm, CPL>(p, stream);
        case kTwoTerms: return launch_narrow<LPR, kTwoTerms, CPL>(p, stream);
        default: return launch_narrow<LPR, kBi, CPL>(p, stream);
    }
}

template <int S, int NC, int MODE>
int launch(const AggParams& p, int threads, size_t smem, cudaStream_t stream) {
    auto kern = aggregate_kernel<S, NC, MODE>;
    LKG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    LKG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) LKG_FAIL(LKG_ERR_UNSUPPORTED, "aggregate kernel does not fit (smem %zu)", smem);
    const int grid = sm_count() * per_sm;
    kern<<<grid, threads, smem, stream>>>(p);
    LKG_LAUNCH_CHECK("aggregate_kernel");
    return LKG_OK;
}

template <int S, int NC>
int dispatch_mode(int mode, const AggParams& p, int threads, size_t smem, cudaStream_t stream) {
    switch (mode) {
        case kOneTerm: return launch<S, NC, kOneTerm>(p, threads, smem, stream);
        case kTwoTerms: return launch<S, NC, kTwoTerms>(p, threads, smem, stream);
        default: return launch<S, NC, kBi>(p, threads, smem, stream);
    }
}

}  // namespace
}  // namespace lkg

using namespace lkg;

extern "C" int lkg_aggregate_workspace_bytes(size_t* bytes) {
    LKG_REQUIRE(bytes != nullptr, "bytes is null");
    *bytes = 256;
    return LKG_OK;
}

extern "C" int lkg_aggregate_fwd(const lkg_graph* g, const float* a_values, const float* ego, int64_t ld_ego,
                                 int32_t d_in, int32_t d_out, const float* pa, const float* pb, const float* p2,
                                 const float* r1, const float* r2, int64_t ld_r, const float* ln_weight,
                                 const float* ln_bias, const float* drop_mask, float* x_out, int64_t ld_x,
                                 float* xn_out, int64_t ld_xn, uint16_t* xn_planes, int64_t ld_planes,
                                 int64_t plane_stride, const float* xn_rec, int64_t local_row_base, void* workspace,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LKG_REQUIRE(g && ego && pb && ln_weight && ln_bias && x_out && workspace, "null argument");
    LKG_REQUIRE(g->nnz == 0 || a_values != nullptr, "a_values is null");
    LKG_REQUIRE(!xn_planes || xn_rec, "xn_planes needs a scale record");
    LKG_REQUIRE(g->row_begin >= 0 && g->row_begin <= g->row_end && g->row_end <= g->n_entities, "bad row range");
    LKG_REQUIRE(d_in > 0 && d_in % 4 == 0, "d_in must be a positive multiple of 4 (got %d)", d_in);
    LKG_REQUIRE(ld_ego % 4 == 0 && aligned16(ego), "ego rows must be 16-byte aligned");
    LKG_REQUIRE(d_out > 0, "d_out must be positive");
    LKG_REQUIRE(!(p2 && pa && pa != pb), "bi-interaction takes one shared sum matrix (pa == pb or pa NULL)");
    if (d_in > 512) LKG_FAIL(LKG_ERR_UNSUPPORTED, "aggregate d_in %d > 512", d_in);
    if (d_out > 64) LKG_FAIL(LKG_ERR_UNSUPPORTED, "aggregate d_out %d > 64", d_out);

    AggParams p{};
    p.g = *g;
    p.a_val = a_values;
    p.ego = ego;
    p.ld_ego = ld_ego;
    p.d_in = d_in;
    p.d_out = d_out;
    p.nvec = d_in / 4;
    int mode;
    if (p2) {
        mode = kBi;
        p.p0 = pb;
        p.p1 = p2;
        p.sum_ego = pa != nullptr;
    } else if (pa && pa != pb) {
        mode = kTwoTerms;
        p.p0 = pa;
        p.p1 = pb;
        p.sum_ego = 0;
    } else {
        mode = kOneTerm;
        p.p0 = pb;
        p.p1 = nullptr;
        p.sum_ego = pa != nullptr;
    }
    p.r1 = r1;
    p.r2 = r2;
    p.ld_r = ld_r;
    p.ln_w = ln_weight;
    p.ln_b = ln_bias;
    p.mask = drop_mask;
    p.x_out = x_out;
    p.ld_x = ld_x;
    p.xn_out = xn_out;
    p.ld_xn = ld_xn;
    p.xn_planes = reinterpret_cast<__half*>(xn_planes);
    p.xn_rec = xn_rec;
    p.ld_planes = ld_planes;
    p.plane_stride = plane_stride;
    p.local_row_base = local_row_base;
    p.counter = static_cast<int*>(workspace);
    // lane c reads sp[d * ps + c]: any stride is conflict free for 32 consecutive channels; pad to a
    // multiple of 4 floats to keep rows 16-byte aligned
    p.p_stride = (d_out + 3) / 4 * 4;
    LKG_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(int), stream));

    // narrow rows: LPR lanes per row (needs 16-byte aligned rows everywhere the kernel moves float4 / 8-byte halves)
    if ((d_in == 16 || d_in == 32 || d_in == 64) && d_out % 4 == 0) {
        const bool al = aligned16(x_out) && ld_x % 4 == 0 && (!xn_out || (aligned16(xn_out) && ld_xn % 4 == 0)) &&
                        (!r1 || (aligned16(r1) && ld_r % 4 == 0)) && (!r2 || (aligned16(r2) && ld_r % 4 == 0)) &&
                        aligned16(ln_weight) && aligned16(ln_bias) && (!drop_mask || aligned16(drop_mask)) &&
                        (!xn_planes || ((reinterpret_cast<uintptr_t>(xn_planes) & 7u) == 0 && ld_planes % 4 == 0 &&
                                        plane_stride % 4 == 0));
        const int chunks = d_out / 4;
        if (al) {
#define LKG_NARROW_CASE(LL, CC) \
    if (d_in == 4 * LL && (chunks + LL - 1) / LL == CC) return dispatch_narrow<LL, CC>(mode, p, stream);
            LKG_NARROW_CASE(4, 1) LKG_NARROW_CASE(4, 2) LKG_NARROW_CASE(4, 3) LKG_NARROW_CASE(4, 4)
            LKG_NARROW_CASE(8, 1) LKG_NARROW_CASE(8, 2)
            LKG_NARROW_CASE(16, 1)
#undef LKG_NARROW_CASE
        }
    }

    const int nt = mode == kOneTerm ? 1 : 2;
    const size_t p_bytes = (size_t)nt * d_in * p.p_stride * sizeof(float);
    const size_t stage_per_warp = (size_t)kRows * nt * d_in * sizeof(float);
    int warps = 16;
    while (warps > 4 && p_bytes + warps * stage_per_warp > 200 * 1024) warps /= 2;
    const size_t smem = p_bytes + warps * stage_per_warp;
    if (smem > 227 * 1024) LKG_FAIL(LKG_ERR_UNSUPPORTED, "aggregate shapes need %zu bytes of shared memory", smem);
    const int threads = warps * 32;
    const int slots = (p.nvec + 31) / 32;
    const int nc = (d_out + 31) / 32;
#define LKG_AGG_CASE(SS, CC) \
    if (slots == SS && nc == CC) return dispatch_mode<SS, CC>(mode, p, threads, smem, stream);
    LKG_AGG_CASE(1, 1) LKG_AGG_CASE(2, 1) LKG_AGG_CASE(3, 1) LKG_AGG_CASE(4, 1)
    LKG_AGG_CASE(1, 2) LKG_AGG_CASE(2, 2) LKG_AGG_CASE(3, 2) LKG_AGG_CASE(4, 2)
#undef LKG_AGG_CASE
    LKG_FAIL(LKG_ERR_UNSUPPORTED, "aggregate: no kernel for d_in %d d_out %d", d_in, d_out);
}
