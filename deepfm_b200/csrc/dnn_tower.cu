// Element-wise / column-reduction half of the DNN tower (reference dnn.py:45-59:
// Linear -> BatchNorm1d -> activation -> Dropout, and the head nn.Linear(., 1) of deepfm.py:36-42).
// The GEMMs are dnn_gemm.cu; here:
//   bn_stats        per-column batch mean / 1/sqrt(var + eps) of the pre-activation (training-mode BatchNorm1d,
//                   biased variance; running statistics updated with the unbiased one like ATen batch_norm)
//   bn_act_fwd      a = dropout(act(gamma * (y - mean) * rstd + beta))        one pass, nothing else materialised
//   bn_act_bwd      dz = da * keep / (1 - p) * act'(z);  dgamma = sum dz * xhat, dbeta = sum dz (column sums), then
//                   dy = gamma * rstd * (dz - dbeta / M - xhat * dgamma / M)   (z, xhat and the dropout mask are
//                   recomputed from y and the counter-based RNG: no mask tensor, no saved activation)
//   col_sum         bias gradient (column sums of dy)
//   head_fwd / bwd  logit = a . w + b  and its gradients
// All column sums are two-stage, fixed-order reductions with fp64 partials: deterministic.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace dfm {
namespace tw {

enum { ACT_RELU = 0, ACT_LEAKY = 1, ACT_GELU = 2, ACT_TANH = 3 };
enum { BN_NONE = 0, BN_BATCH = 1, BN_FIXED = 2 };

constexpr int SLAB_ROWS = 256;       // rows per partial
constexpr int CT = 32, RT = 8;       // block = 32 columns x 8 row lanes

__device__ __forceinline__ float act_f(float z, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(z, 0.f);
        case ACT_LEAKY: return z > 0.f ? z : 0.01f * z;
        case ACT_GELU: return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
        default: return tanhf(z);
    }
}
__device__ __forceinline__ float act_df(float z, int act) {
    switch (act) {
        case ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case ACT_LEAKY: return z > 0.f ? 1.f : 0.01f;
        case ACT_GELU: {
            const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
            return cdf + z * 0.39894228040143267794f * expf(-0.5f * z * z);
        }
        default: { const float t = tanhf(z); return 1.f - t * t; }
    }
}
// counter-based uniform in [0, 1): splitmix64 of (seed, element index) -- the backward regenerates the same mask
// One splitmix64 per FOUR consecutive elements (16 bits each: the keep decision has a resolution of 1 / 65536): the three
// 64-bit multiplies per element were a third of the instructions of every element-wise pass.
__device__ __forceinline__ unsigned long long mix64(unsigned long long seed, unsigned long long idx4) {
    unsigned long long z = seed + (idx4 + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01_of(unsigned long long z, int field) {
    return (float)((unsigned)(z >> (16 * field)) & 0xffffu) * (1.0f / 65536.0f);
}
__device__ __forceinline__ float u01(unsigned long long seed, unsigned long long idx) {
    return u01_of(mix64(seed, idx >> 2), (int)(idx & 3ull));
}

struct Ew {                      // element-wise context of one layer
    const float* y;              // (M, C) pre-activation (Linear output)
    const float* mean; const float* rstd; const float* gamma; const float* beta;   // BN (null when bn == BN_NONE)
    long long M; int C;
    int bn, act;
    float p, keep_scale;         // dropout probability, 1 / (1 - p)
    unsigned long long seed;
};

__device__ __forceinline__ float pre_act(const Ew& e, float y, int c, float& xhat) {
    if (e.bn == BN_NONE) { xhat = 0.f; return y; }
    xhat = (y - __ldg(e.mean + c)) * __ldg(e.rstd + c);
    return fmaf(__ldg(e.gamma + c), xhat, __ldg(e.beta + c));
}

// ---- column sums: partial[slab][k][c] (double), k = 0, 1
// OP 0: (sum y, sum y^2);  OP 1: (sum dz * xhat, sum dz) with dz recomputed;  OP 2: (sum x, -)
// A block owns SLAB_ROWS rows x (32 * VC) columns: thread (tx, ty) walks rows ty, ty + RT, ... of VC consecutive
// columns (one 128-bit load when VC = 4), UR rows in flight; fp32 terms, fp64 running sums.
template <int OP>
__device__ __forceinline__ void colsum_term(const Ew& e, float y, float d, float u, int c, double& a0, double& a1) {
    if (OP == 0) {
        a0 += (double)y; a1 += (double)y * (double)y;
    } else if (OP == 1) {
        float xhat;
        const float z = pre_act(e, y, c, xhat);
        float g = d * act_df(z, e.act);
        if (e.p > 0.f) g = u >= e.p ? g * e.keep_scale : 0.f;
        a0 += (double)g * (double)xhat; a1 += (double)g;
    } else {
        a0 += (double)d;
    }
}

template <int OP, int VC>
__global__ void __launch_bounds__(CT * RT)
colsum_partial_kernel(const __grid_constant__ Ew e, const float* __restrict__ da, double* __restrict__ partial) {
    __shared__ double s0[RT][CT * VC], s1[RT][CT * VC];
    constexpr int UR = 4;
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    const int c0 = (blockIdx.y * CT + tx) * VC;
    const long long r0 = (long long)blockIdx.x * SLAB_ROWS;
    const long long r1 = r0 + SLAB_ROWS < e.M ? r0 + SLAB_ROWS : e.M;
    double a0[VC], a1[VC];
#pragma unroll
    for (int v = 0; v < VC; ++v) { a0[v] = 0.0; a1[v] = 0.0; }
    if (c0 < e.C) {
        for (long long r = r0 + ty; r < r1; r += RT * UR) {
            float yv[UR][VC], dv[UR][VC];
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const long long rr = r + (long long)u * RT;
                const size_t i = (size_t)(rr < r1 ? rr : r) * e.C + c0;
                if (VC == 4) {
                    if (OP != 2) { const float4 t = __ldg(reinterpret_cast<const float4*>(e.y + i)); yv[u][0] = t.x; yv[u][1] = t.y; yv[u][2] = t.z; yv[u][3] = t.w; }
                    if (OP != 0) { const float4 t = __ldg(reinterpret_cast<const float4*>(da + i)); dv[u][0] = t.x; dv[u][1] = t.y; dv[u][2] = t.z; dv[u][3] = t.w; }
                } else {
                    if (OP != 2) yv[u][0] = __ldg(e.y + i);
                    if (OP != 0) dv[u][0] = __ldg(da + i);
                }
            }
#pragma unroll
            for (int u = 0; u < UR; ++u) {
                const long long rr = r + (long long)u * RT;
                if (rr >= r1) break;
                const size_t i0 = (size_t)rr * e.C + c0;
                unsigned long long zr = 0ull;
                if (OP == 1 && e.p > 0.f && VC == 4) zr = mix64(e.seed, i0 >> 2);       // c0 and C are multiples of 4 here
#pragma unroll
                for (int v = 0; v < VC; ++v) {
                    float uu = 0.f;
                    if (OP == 1 && e.p > 0.f) uu = VC == 4 ? u01_of(zr, v) : u01(e.seed, i0 + v);
                    colsum_term<OP>(e, OP != 2 ? yv[u][v] : 0.f, OP != 0 ? dv[u][v] : 0.f, uu, c0 + v, a0[v], a1[v]);
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VC; ++v) { s0[ty][tx * VC + v] = a0[v]; s1[ty][tx * VC + v] = a1[v]; }
    __syncthreads();
    for (int o = threadIdx.x; o < CT * VC; o += CT * RT) {
        const int c = blockIdx.y * CT * VC + o;
        if (c >= e.C) continue;
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int q = 0; q < RT; ++q) { t0 += s0[q][o]; t1 += s1[q][o]; }
        partial[((size_t)blockIdx.x * 2 + 0) * e.C + c] = t0;
        partial[((size_t)blockIdx.x * 2 + 1) * e.C + c] = t1;
    }
}

// second stage: (s, q)[c] = sum over slabs in slab order; a block owns 32 columns, 8 lanes stride over the slabs
__device__ __forceinline__ void colsum_second_stage(const double* __restrict__ partial, int n_slab, int C, int stride_c,
                                                    double& s, double& q, bool& owner, int& c) {
    __shared__ double f0[RT][CT], f1[RT][CT];
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    c = blockIdx.x * CT + tx;
    double a = 0.0, b = 0.0;
    if (c < C) {
        for (int sl = ty; sl < n_slab; sl += RT) {
            a += partial[((size_t)sl * 2) * stride_c + c];
            b += partial[((size_t)sl * 2 + 1) * stride_c + c];
        }
    }
    f0[ty][tx] = a; f1[ty][tx] = b;
    __syncthreads();
    owner = ty == 0 && c < C;
    s = 0.0; q = 0.0;
    if (owner) {
#pragma unroll
        for (int k = 0; k < RT; ++k) { s += f0[k][tx]; q += f1[k][tx]; }
    }
}

// finalize OP 0: mean, rstd (+ running statistics, momentum update with the unbiased variance)
__global__ void __launch_bounds__(CT * RT)
bn_stats_final_kernel(const double* __restrict__ partial, int n_slab, long long M, int C, float eps,
                      float* __restrict__ mean, float* __restrict__ rstd,
                      float* __restrict__ run_mean, float* __restrict__ run_var, float momentum) {
    double s, q; bool owner; int c;
    colsum_second_stage(partial, n_slab, C, C, s, q, owner, c);
    if (!owner) return;
    const double mu = s / (double)M;
    double var = q / (double)M - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (run_mean) {
        const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
        run_mean[c] = (float)((1.0 - momentum) * run_mean[c] + momentum * mu);
        run_var[c] = (float)((1.0 - momentum) * run_var[c] + momentum * unb);
    }
}

// finalize OP 1 / 2: two float vectors (second optional)
__global__ void __launch_bounds__(CT * RT)
colsum_final_kernel(const double* __restrict__ partial, int n_slab, int C, float* __restrict__ out0, float* __restrict__ out1) {
    double s, q; bool owner; int c;
    colsum_second_stage(partial, n_slab, C, C, s, q, owner, c);
    if (!owner) return;
    if (out0) out0[c] = (float)s;
    if (out1) out1[c] = (float)q;
}

// element-wise passes: VE consecutive elements per thread (one 128-bit access when the width is a multiple of 4)
template <int VE>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const __grid_constant__ Ew e, float* __restrict__ out) {
    const long long n = e.M * e.C / VE;
    const bool small = n < 0x7fffffffLL;          // 32-bit column arithmetic (a 64-bit modulo per vector otherwise)
    const unsigned cvec = (unsigned)(e.C / VE);
    for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < n; iv += (long long)gridDim.x * blockDim.x) {
        const long long i = iv * VE;
        const int c = small ? (int)(((unsigned)iv % cvec) * VE) : (int)(i % e.C);
        float y[VE], a[VE];
        if (VE == 4) { const float4 t = __ldcs(reinterpret_cast<const float4*>(e.y + i)); y[0] = t.x; y[1] = t.y; y[2] = t.z; y[3] = t.w; }
        else y[0] = __ldcs(e.y + i);
        unsigned long long zr = 0ull;
        if (e.p > 0.f && VE == 4) zr = mix64(e.seed, (unsigned long long)iv);
#pragma unroll
        for (int v = 0; v < VE; ++v) {
            float xhat;
            a[v] = act_f(pre_act(e, y[v], c + v, xhat), e.act);
            if (e.p > 0.f) a[v] = (VE == 4 ? u01_of(zr, v) : u01(e.seed, (unsigned long long)(i + v))) >= e.p ? a[v] * e.keep_scale : 0.f;
        }
        if (VE == 4) *reinterpret_cast<float4*>(out + i) = make_float4(a[0], a[1], a[2], a[3]);
        else out[i] = a[0];
    }
}

// dy = gamma * rstd * (dz - dbeta / M - xhat * dgamma / M)      (BN_BATCH)
//    = gamma * rstd * dz                                        (BN_FIXED)      = dz (BN_NONE)
template <int VE>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const __grid_constant__ Ew e, const float* __restrict__ da, const float* __restrict__ dgamma,
                        const float* __restrict__ dbeta, float* __restrict__ dy) {
    const long long n = e.M * e.C / VE;
    const bool small = n < 0x7fffffffLL;
    const unsigned cvec = (unsigned)(e.C / VE);
    const float inv_m = 1.f / (float)e.M;
    for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < n; iv += (long long)gridDim.x * blockDim.x) {
        const long long i = iv * VE;
        const int c = small ? (int)(((unsigned)iv % cvec) * VE) : (int)(i % e.C);
        float y[VE], d[VE], g[VE];
        unsigned long long zr = 0ull;
        if (e.p > 0.f && VE == 4) zr = mix64(e.seed, (unsigned long long)iv);
        if (VE == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(e.y + i)); y[0] = t.x; y[1] = t.y; y[2] = t.z; y[3] = t.w;
            const float4 u = __ldcs(reinterpret_cast<const float4*>(da + i)); d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
        } else { y[0] = __ldcs(e.y + i); d[0] = __ldcs(da + i); }
#pragma unroll
        for (int v = 0; v < VE; ++v) {
            float xhat;
            const float z = pre_act(e, y[v], c + v, xhat);
            g[v] = d[v] * act_df(z, e.act);
            if (e.p > 0.f) g[v] = (VE == 4 ? u01_of(zr, v) : u01(e.seed, (unsigned long long)(i + v))) >= e.p ? g[v] * e.keep_scale : 0.f;
            if (e.bn == BN_BATCH)
                g[v] = __ldg(e.gamma + c + v) * __ldg(e.rstd + c + v) * (g[v] - __ldg(dbeta + c + v) * inv_m - xhat * __ldg(dgamma + c + v) * inv_m);
            else if (e.bn == BN_FIXED)
                g[v] = __ldg(e.gamma + c + v) * __ldg(e.rstd + c + v) * g[v];
        }
        if (VE == 4) *reinterpret_cast<float4*>(dy + i) = make_float4(g[0], g[1], g[2], g[3]);
        else dy[i] = g[0];
    }
}

// ---- head: logit[m] = a[m, :] . w + b   (one warp per row)
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ b, long long M, int C,
                float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= M) return;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(__ldg(a + (size_t)row * C + c), __ldg(w + c), acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc + (b ? __ldg(b) : 0.f);
}

// da[m, c] = g[m] * w[c];   partial sums of dw[c] = sum_m g[m] a[m, c] and db = sum_m g[m] (column C of the partial)
__global__ void __launch_bounds__(CT * RT)
head_bwd_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ g, long long M, int C,
                float* __restrict__ da, double* __restrict__ partial) {
    __shared__ double s0[RT][CT];
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    const int c = blockIdx.y * CT + tx;               // column C is the bias pseudo-column
    const long long r0 = (long long)blockIdx.x * SLAB_ROWS;
    const long long r1 = r0 + SLAB_ROWS < M ? r0 + SLAB_ROWS : M;
    double acc = 0.0;
    if (c <= C) {
        const float wc = c < C ? __ldg(w + c) : 0.f;
        for (long long r = r0 + ty; r < r1; r += RT) {
            const float gr = __ldg(g + r);
            if (c < C) {
                acc += (double)gr * (double)__ldg(a + (size_t)r * C + c);
                if (da) da[(size_t)r * C + c] = gr * wc;
            } else {
                acc += (double)gr;
            }
        }
    }
    s0[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c <= C) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < RT; ++q) t += s0[q][tx];
        partial[((size_t)blockIdx.x * 2) * (C + 1) + c] = t;
        partial[((size_t)blockIdx.x * 2 + 1) * (C + 1) + c] = 0.0;
    }
}

__global__ void __launch_bounds__(CT * RT)
head_final_kernel(const double* __restrict__ partial, int n_slab, int C, float* __restrict__ dw, float* __restrict__ db) {
    double s, q; bool owner; int c;
    colsum_second_stage(partial, n_slab, C + 1, C + 1, s, q, owner, c);
    if (!owner) return;
    if (c < C) dw[c] = (float)s;
    else if (db) db[0] = (float)s;
}

static inline int n_slabs(long long M) { return (int)ceil_div(M > 0 ? M : 1, SLAB_ROWS); }

static int fill_ew(Ew& e, const float* y, long long M, int C, int bn, int act, const float* mean, const float* rstd,
                   const float* gamma, const float* beta, float p, unsigned long long seed, const char* who) {
    DFM_REQUIRE(y && M >= 0 && C > 0, DFM_ERR_INVALID, "%s: bad argument", who);
    DFM_REQUIRE(bn >= BN_NONE && bn <= BN_FIXED && act >= ACT_RELU && act <= ACT_TANH, DFM_ERR_INVALID, "%s: unknown bn / activation code", who);
    DFM_REQUIRE(bn == BN_NONE || (mean && rstd && gamma && beta), DFM_ERR_INVALID, "%s: BatchNorm tensors missing", who);
    DFM_REQUIRE(p >= 0.f && p < 1.f, DFM_ERR_INVALID, "%s: dropout p must be in [0, 1)", who);
    memset(&e, 0, sizeof(e));
    e.y = y; e.M = M; e.C = C; e.bn = bn; e.act = act; e.mean = mean; e.rstd = rstd; e.gamma = gamma; e.beta = beta;
    e.p = p; e.keep_scale = 1.f / (1.f - p); e.seed = seed;
    return DFM_OK;
}

static inline bool vec4(int C, const void* p) { return C % 4 == 0 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline unsigned ew_blocks(long long n) {
    long long b = ceil_div(n > 0 ? n : 1, 256 * 4);
    const long long cap = 8LL * sm_count();
    return (unsigned)(b < cap ? b : cap);
}

}  // namespace tw
}  // namespace dfm

using namespace dfm;
using namespace dfm::tw;

extern "C" {

size_t dfm_tower_workspace_bytes(int64_t M, int C) {
    return (size_t)n_slabs(M) * 2 * (size_t)(C + 1) * sizeof(double) + 256;
}

int dfm_bn_stats(const float* y, int64_t M, int C, float eps, float* mean, float* rstd, float* running_mean,
                 float* running_var, float momentum, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(y && mean && rstd && workspace && M > 0 && C > 0, DFM_ERR_INVALID, "dfm_bn_stats: bad argument");
    DFM_REQUIRE(workspace_bytes >= dfm_tower_workspace_bytes(M, C), DFM_ERR_WORKSPACE, "dfm_bn_stats: workspace too small");
    DFM_REQUIRE((running_mean == nullptr) == (running_var == nullptr), DFM_ERR_INVALID, "dfm_bn_stats: running stats come in pairs");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Ew e;
    memset(&e, 0, sizeof(e));
    e.y = y; e.M = M; e.C = C;
    const int ns = n_slabs(M);
    double* partial = static_cast<double*>(workspace);
    if (vec4(C, y)) colsum_partial_kernel<0, 4><<<dim3(ns, (unsigned)ceil_div(C, CT * 4)), CT * RT, 0, st>>>(e, nullptr, partial);
    else colsum_partial_kernel<0, 1><<<dim3(ns, (unsigned)ceil_div(C, CT)), CT * RT, 0, st>>>(e, nullptr, partial);
    bn_stats_final_kernel<<<(unsigned)ceil_div(C, CT), CT * RT, 0, st>>>(partial, ns, M, C, eps, mean, rstd, running_mean, running_var, momentum);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_bn_act_fwd(const float* y, int64_t M, int C, int bn, int act, const float* mean, const float* rstd,
                   const float* gamma, const float* beta, float drop_p, uint64_t seed, float* out, void* stream) {
    Ew e;
    int rc = fill_ew(e, y, M, C, bn, act, mean, rstd, gamma, beta, drop_p, seed, "dfm_bn_act_fwd");
    if (rc) return rc;
    DFM_REQUIRE(out, DFM_ERR_INVALID, "dfm_bn_act_fwd: null output");
    if (M == 0) return DFM_OK;
    if (vec4(C, y) && vec4(C, out)) bn_act_fwd_kernel<4><<<ew_blocks(M * C / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(e, out);
    else bn_act_fwd_kernel<1><<<ew_blocks(M * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(e, out);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_bn_act_bwd(const float* da, const float* y, int64_t M, int C, int bn, int act, const float* mean, const float* rstd,
                   const float* gamma, const float* beta, float drop_p, uint64_t seed, float* dy, float* dgamma, float* dbeta,
                   float* dbias, void* workspace, size_t workspace_bytes, void* stream) {
    Ew e;
    int rc = fill_ew(e, y, M, C, bn, act, mean, rstd, gamma, beta, drop_p, seed, "dfm_bn_act_bwd");
    if (rc) return rc;
    DFM_REQUIRE(da && dy && workspace, DFM_ERR_INVALID, "dfm_bn_act_bwd: null tensor");
    DFM_REQUIRE(bn == BN_NONE || (dgamma && dbeta), DFM_ERR_INVALID, "dfm_bn_act_bwd: BatchNorm gradient outputs missing");
    DFM_REQUIRE(workspace_bytes >= dfm_tower_workspace_bytes(M, C), DFM_ERR_WORKSPACE, "dfm_bn_act_bwd: workspace too small");
    if (M == 0) return DFM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ns = n_slabs(M);
    double* partial = static_cast<double*>(workspace);
    const bool v4 = vec4(C, y) && vec4(C, da) && vec4(C, dy);
    const dim3 grid(ns, (unsigned)ceil_div(C, v4 ? CT * 4 : CT));
    if (bn != BN_NONE) {
        if (v4) colsum_partial_kernel<1, 4><<<grid, CT * RT, 0, st>>>(e, da, partial);
        else colsum_partial_kernel<1, 1><<<grid, CT * RT, 0, st>>>(e, da, partial);
        colsum_final_kernel<<<(unsigned)ceil_div(C, CT), CT * RT, 0, st>>>(partial, ns, C, dgamma, dbeta);
    }
    if (v4) bn_act_bwd_apply_kernel<4><<<ew_blocks(M * C / 4), 256, 0, st>>>(e, da, dgamma, dbeta, dy);
    else bn_act_bwd_apply_kernel<1><<<ew_blocks(M * C), 256, 0, st>>>(e, da, dgamma, dbeta, dy);
    if (dbias) {     // Linear bias gradient: column sums of dy (analytically zero behind a training-mode BatchNorm)
        Ew e2;
        memset(&e2, 0, sizeof(e2));
        e2.M = M; e2.C = C;
        if (v4) colsum_partial_kernel<2, 4><<<grid, CT * RT, 0, st>>>(e2, dy, partial);
        else colsum_partial_kernel<2, 1><<<grid, CT * RT, 0, st>>>(e2, dy, partial);
        colsum_final_kernel<<<(unsigned)ceil_div(C, CT), CT * RT, 0, st>>>(partial, ns, C, dbias, nullptr);
    }
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_head_fwd(const float* a, const float* w, const float* b, int64_t M, int C, float* out, void* stream) {
    DFM_REQUIRE(a && w && out && M >= 0 && C > 0, DFM_ERR_INVALID, "dfm_head_fwd: bad argument");
    if (M == 0) return DFM_OK;
    head_fwd_kernel<<<(unsigned)ceil_div(M * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, w, b, M, C, out);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_head_bwd(const float* a, const float* w, const float* g, int64_t M, int C, float* da, float* dw, float* db,
                 void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(a && w && g && dw && workspace && M > 0 && C > 0, DFM_ERR_INVALID, "dfm_head_bwd: bad argument");
    DFM_REQUIRE(workspace_bytes >= dfm_tower_workspace_bytes(M, C), DFM_ERR_WORKSPACE, "dfm_head_bwd: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ns = n_slabs(M);
    double* partial = static_cast<double*>(workspace);
    head_bwd_kernel<<<dim3(ns, (unsigned)ceil_div(C + 1, CT)), CT * RT, 0, st>>>(a, w, g, M, C, da, partial);
    // dw = columns [0, C), db = column C of the (C + 1)-wide partials
    head_final_kernel<<<(unsigned)ceil_div(C + 1, CT), CT * RT, 0, st>>>(partial, ns, C, dw, db);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
