// Row-sparse Adam on the embedding tables (reference step body trainer.py:232-237:
// clip_grad_norm_(1.0) then torch.optim.Adam(lr 1e-3), trainer.py:67-78).
//
// K2's row-sparse output is consumed as it is: sorted keys (global row = row_base[f] + id, PAD last) and, at the
// first sorted position of every run of equal keys, that row's summed gradient (row_grad2 (N, d_max), row_grad1 (N)).
// One lane group per sorted position; only segment heads do work: m, v and w of that ONE table row (and of its
// first-order scalar) are updated in place -- torch.optim.Adam's update restricted to the touched rows ("lazy"
// moments: untouched rows keep m, v and w; the documented deviation from the reference's dense Adam, whose
// untouched rows still decay through their stale moments).  No compaction, no host sync.
//   rows_sumsq: sum of squares of the head rows (the table part of the global gradient norm), per-block partial
//   sums added in block order: deterministic.
#include <math.h>
#include <string.h>

#include "plan.cuh"

namespace dfm {

struct AdamField {
    float *w2, *w1, *m2, *m1, *v2, *v1;
    unsigned row_base, row_end;
    int dim;
};
struct AdamArgs {
    AdamField f[MAX_FIELDS];
    int n;                 // table fields
    int tdim;              // row stride of row_grad2
    unsigned pad;
    float lr, b1, b2, eps, bc1, bc2_sqrt;   // bias corrections 1 - b1^t, sqrt(1 - b2^t)
    const float* clip;     // device scalar multiplying every gradient (global-norm clip), or null
};

__device__ __forceinline__ int adam_field_of(const AdamField* t, int n, uint32_t key) {
    int f = 0;
    for (int i = 1; i < n; ++i) if (key >= t[i].row_base) f = i;
    return f;
}

__device__ __forceinline__ float adam_update(float g, float& m, float& v, float w, const AdamArgs& a) {
    m = a.b1 * m + (1.f - a.b1) * g;
    v = a.b2 * v + (1.f - a.b2) * g * g;
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;          // torch: (sqrt(v) / sqrt(bc2)) + eps
    return w - (a.lr / a.bc1) * (m / denom);
}

template <int V>
__global__ void __launch_bounds__(256)
adam_rows_kernel(const __grid_constant__ AdamArgs a, long long N, const uint32_t* __restrict__ skeys,
                 const float* __restrict__ rg2, const float* __restrict__ rg1, int G, double* __restrict__ ssq_partial) {
    __shared__ AdamField t[MAX_FIELDS];
    __shared__ double s_red[8];
    double dsq = 0.0;      // sum over this thread's elements of w_new^2 - w_old^2 (incremental ||W||^2)
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) t[i] = a.f[i];
    __syncthreads();
    const int gpb = blockDim.x / G, gl = threadIdx.x / G, j = threadIdx.x - gl * G;
    const float clip = a.clip ? __ldg(a.clip) : 1.f;
    for (long long p = (long long)blockIdx.x * gpb + gl; p < N; p += (long long)gridDim.x * gpb) {
        const uint32_t key = __ldg(skeys + p);
        if (key == a.pad) break;                                   // PAD keys sort last
        if (p > 0 && __ldg(skeys + p - 1) == key) continue;        // not the head of its segment
        const AdamField& fd = t[adam_field_of(t, a.n, key)];
        const size_t row = key - fd.row_base;
        if (j < fd.dim / V) {
            const size_t o = row * fd.dim + j * V;
            VecF<V> g = vload<V>(rg2 + (size_t)p * a.tdim + j * V);
            VecF<V> w = vload<V>(fd.w2 + o), m = vload<V>(fd.m2 + o), v = vload<V>(fd.v2 + o);
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const float wo = w.v[q];
                w.v[q] = adam_update(g.v[q] * clip, m.v[q], v.v[q], wo, a);
                if (ssq_partial) dsq += (double)w.v[q] * (double)w.v[q] - (double)wo * (double)wo;
            }
            vstore<V>(fd.w2 + o, w); vstore<V>(fd.m2 + o, m); vstore<V>(fd.v2 + o, v);
        }
        if (j == 0) {
            float m = fd.m1[row], v = fd.v1[row];
            const float wo = fd.w1[row];
            const float wn = adam_update(__ldg(rg1 + p) * clip, m, v, wo, a);
            fd.w1[row] = wn;
            fd.m1[row] = m; fd.v1[row] = v;
            if (ssq_partial) dsq += (double)wn * (double)wn - (double)wo * (double)wo;
        }
    }
    if (ssq_partial) {      // per-block partial, fixed order inside the block; blocks are added in block order
        for (int o = 16; o > 0; o >>= 1) dsq += __shfl_xor_sync(0xffffffffu, dsq, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dsq;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += s_red[w];
            ssq_partial[blockIdx.x] = s;
        }
    }
}

__global__ void ssq_apply_kernel(const double* __restrict__ partial, int n, double* __restrict__ acc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += partial[i];
        acc[0] += s;
    }
}

struct RowDims {            // table fields in key order: only dim[f] floats of a row_grad2 row are defined
    unsigned row_base[MAX_FIELDS];
    int dim[MAX_FIELDS];
    int n;
};

__global__ void __launch_bounds__(256)
rows_sumsq_kernel(long long N, const uint32_t* __restrict__ skeys, unsigned pad, const float* __restrict__ rg2,
                  const float* __restrict__ rg1, int tdim, float* __restrict__ partial, const __grid_constant__ RowDims rd) {
    __shared__ float red[8];
    float acc = 0.f;
    const long long per = (N + gridDim.x - 1) / gridDim.x;        // contiguous positions per block
    const long long lo = (long long)blockIdx.x * per, hi = lo + per < N ? lo + per : N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long p = lo + warp; p < hi; p += 8) {               // one warp per position, lanes over the row
        const uint32_t key = __ldg(skeys + p);
        if (key == pad || (p > 0 && __ldg(skeys + p - 1) == key)) continue;
        int fi = 0;
        for (int q = 1; q < rd.n; ++q) if (key >= rd.row_base[q]) fi = q;
        const int dimf = rd.dim[fi];                              // K2 writes dim[f] floats of the row, the tail is undefined
        for (int c = lane; c < dimf; c += 32) { const float g = __ldg(rg2 + (size_t)p * tdim + c); acc = fmaf(g, g, acc); }
        if (lane == 0) { const float g = __ldg(rg1 + p); acc = fmaf(g, g, acc); }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void rows_sumsq_final_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += (double)partial[i];
        out[0] = (float)s;
    }
}

}  // namespace dfm

using namespace dfm;

extern "C" {

size_t dfm_adam_rows_workspace_bytes(void) { return (size_t)8 * 8 * 256 * sizeof(double); }   /* >= 8 * SMs blocks */

int dfm_adam_rows(const dfm_plan* plan, int64_t n_sorted, const uint32_t* sorted_keys, const float* row_grad2,
                  const float* row_grad1, float* const* params, float* const* exp_avg, float* const* exp_avg_sq,
                  float lr, float beta1, float beta2, float eps, int64_t step, const float* clip_scale,
                  double* ssq_acc, void* ssq_workspace, size_t ssq_workspace_bytes, void* stream) {
    DFM_REQUIRE(plan && params && exp_avg && exp_avg_sq && step >= 1, DFM_ERR_INVALID, "dfm_adam_rows: bad argument");
    DFM_REQUIRE(!ssq_acc || (ssq_workspace && ssq_workspace_bytes >= dfm_adam_rows_workspace_bytes()), DFM_ERR_WORKSPACE,
                "dfm_adam_rows: ssq_acc needs a workspace of dfm_adam_rows_workspace_bytes()");
    if (n_sorted <= 0 || plan->S == 0) return DFM_OK;
    DFM_REQUIRE(sorted_keys && row_grad2 && row_grad1, DFM_ERR_INVALID, "dfm_adam_rows: null tensor");
    AdamArgs* a = new AdamArgs;
    struct Gd { AdamArgs* p; ~Gd() { delete p; } } gd{a};
    memset(a, 0, sizeof(*a));
    bool aligned = (reinterpret_cast<uintptr_t>(row_grad2) & 15u) == 0;
    for (int f = 0; f < plan->n_fields; ++f) {
        if (plan->kind[f] == DFM_DENSE || plan->foreign[f]) continue;
        AdamField& af = a->f[a->n++];
        af.w2 = params[5 * f]; af.w1 = params[5 * f + 2];
        af.m2 = exp_avg[5 * f]; af.m1 = exp_avg[5 * f + 2];
        af.v2 = exp_avg_sq[5 * f]; af.v1 = exp_avg_sq[5 * f + 2];
        DFM_REQUIRE(af.w2 && af.w1 && af.m2 && af.m1 && af.v2 && af.v1, DFM_ERR_INVALID, "dfm_adam_rows: field %d has a null tensor", f);
        af.row_base = (unsigned)plan->row_base[f]; af.row_end = (unsigned)plan->row_base[f + 1]; af.dim = plan->dim[f];
        aligned = aligned && ((reinterpret_cast<uintptr_t>(af.w2) | reinterpret_cast<uintptr_t>(af.m2) | reinterpret_cast<uintptr_t>(af.v2)) & 15u) == 0;
    }
    a->tdim = plan->max_tdim; a->pad = (unsigned)plan->total_rows;
    a->lr = lr; a->b1 = beta1; a->b2 = beta2; a->eps = eps;
    a->bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a->bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a->clip = clip_scale;
    const int V = (plan->vec == 4 && aligned) ? 4 : 1;
    const int lanes = plan->max_tdim / V;
    DFM_REQUIRE(lanes >= 1 && lanes <= 32, DFM_ERR_UNSUPPORTED, "dfm_adam_rows: table dim %d too wide", plan->max_tdim);
    const int G = next_pow2(lanes);
    long long blocks = ceil_div(n_sorted, 256 / G);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    if (blocks > 8 * 256) blocks = 8 * 256;                      // dfm_adam_rows_workspace_bytes() partial slots
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* part = ssq_acc ? static_cast<double*>(ssq_workspace) : nullptr;
    if (V == 4) adam_rows_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*a, n_sorted, sorted_keys, row_grad2, row_grad1, G, part);
    else adam_rows_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*a, n_sorted, sorted_keys, row_grad2, row_grad1, G, part);
    if (ssq_acc) ssq_apply_kernel<<<1, 32, 0, st>>>(part, (int)blocks, ssq_acc);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

size_t dfm_rows_sumsq_workspace_bytes(void) { return 1024 * sizeof(float); }

int dfm_rows_sumsq(const dfm_plan* plan, int64_t n_sorted, const uint32_t* sorted_keys, const float* row_grad2,
                   const float* row_grad1, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(plan && out && workspace && workspace_bytes >= 1024 * sizeof(float), DFM_ERR_INVALID, "dfm_rows_sumsq: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* partial = static_cast<float*>(workspace);
    int blocks = 0;
    if (n_sorted > 0 && plan->S > 0) {
        DFM_REQUIRE(sorted_keys && row_grad2 && row_grad1, DFM_ERR_INVALID, "dfm_rows_sumsq: null tensor");
        blocks = (int)(ceil_div(n_sorted, 64) < 1024 ? ceil_div(n_sorted, 64) : 1024);
        RowDims rd;
        memset(&rd, 0, sizeof(rd));
        for (int f = 0; f < plan->n_fields; ++f) {
            if (plan->kind[f] == DFM_DENSE || plan->foreign[f]) continue;
            rd.row_base[rd.n] = (unsigned)plan->row_base[f]; rd.dim[rd.n] = plan->dim[f]; ++rd.n;
        }
        rows_sumsq_kernel<<<blocks, 256, 0, st>>>(n_sorted, sorted_keys, (unsigned)plan->total_rows, row_grad2, row_grad1,
                                                  plan->max_tdim, partial, rd);
    }
    rows_sumsq_final_kernel<<<1, 32, 0, st>>>(partial, blocks, out);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
