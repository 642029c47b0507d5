// Stand-alone FMInteraction (fm.py:18-23) for tensors that do not come out of K1, and the
// L2 penalty value lambda * sum_p ||p||^2 of BaseCTRModel.get_l2_reg_loss (base.py:78-83).
#include "common.cuh"

namespace dfm {

// One warp per sample: lanes stride over the D dims, loop over fields; 0.5 * sum_d (S^2 - Q).
__global__ void __launch_bounds__(256)
fm_fwd_kernel(const float* __restrict__ e, long long B, int F, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const float* eb = e + (size_t)b * F * D;
    float part = 0.f;
    for (int d = lane; d < D; d += 32) {
        float s = 0.f, q = 0.f;
        for (int f = 0; f < F; ++f) {
            const float x = __ldg(eb + (size_t)f * D + d);
            s += x;
            q += __fmul_rn(x, x);
        }
        part += __fmul_rn(s, s) - q;   // no FMA contraction: a single field must give exactly 0
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) out[b] = 0.5f * part;
}

// d/de[b,f,d] = g[b] * (S[b,d] - e[b,f,d])
__global__ void __launch_bounds__(256)
fm_bwd_kernel(const float* __restrict__ e, const float* __restrict__ g, long long B, int F, int D,
              float* __restrict__ ge) {
    const int lane = threadIdx.x & 31;
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const float* eb = e + (size_t)b * F * D;
    float* gb = ge + (size_t)b * F * D;
    const float gv = __ldg(g + b);
    for (int d = lane; d < D; d += 32) {
        float s = 0.f;
        for (int f = 0; f < F; ++f) s += __ldg(eb + (size_t)f * D + d);
        for (int f = 0; f < F; ++f) gb[(size_t)f * D + d] = gv * (s - __ldg(eb + (size_t)f * D + d));
    }
}

// ---- sum of squares over a tensor list: fixed grid, fixed order => deterministic ------------
constexpr int SS_BLOCKS = 1024;
constexpr int SS_MAXT = 64;
struct SumsqArgs {
    const float* p[SS_MAXT];
    long long n[SS_MAXT];
    int nt;
};

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const __grid_constant__ SumsqArgs a, float* __restrict__ partial, int accumulate) {
    __shared__ float red[8];
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int t = 0; t < a.nt; ++t) {
        const float* p = a.p[t];
        const long long n = a.n[t];
        if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
            const long long n4 = n >> 2;
            for (long long i = t0; i < n4; i += stride) {
                const float4 w = __ldcs(reinterpret_cast<const float4*>(p) + i);
                acc += (w.x * w.x + w.y * w.y) + (w.z * w.z + w.w * w.w);
            }
            for (long long i = (n4 << 2) + t0; i < n; i += stride) acc = fmaf(p[i], p[i], acc);
        } else {
            for (long long i = t0; i < n; i += stride) acc = fmaf(p[i], p[i], acc);
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        partial[blockIdx.x] = accumulate ? partial[blockIdx.x] + s : s;
    }
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
    __shared__ double red[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)partial[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        out[0] = (float)(s * (double)scale);
    }
}

// double-precision variant of the final stage: acc[0] = sum of the per-block partials (no scale), the exact
// starting point of the incrementally maintained ||W||^2 (dfm_adam_rows adds sum(new^2 - old^2) of touched rows)
__global__ void __launch_bounds__(256)
sumsq_final_f64_kernel(const float* __restrict__ partial, int n, double* __restrict__ acc) {
    __shared__ double red[8];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += (double)partial[i];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        acc[0] = s;
    }
}

__global__ void l2_combine_kernel(const double* __restrict__ a, const double* __restrict__ b, float lam,
                                  float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)((double)lam * (a[0] + (b ? b[0] : 0.0)));
}

__global__ void axpy_kernel(const float* __restrict__ p, long long n, float coef,
                            const float* __restrict__ scale_dev, float* __restrict__ g, int accumulate) {
    const float c = coef * (scale_dev ? __ldg(scale_dev) : 1.f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        g[i] = accumulate ? fmaf(c, p[i], g[i]) : c * p[i];
}

}  // namespace dfm

using namespace dfm;

extern "C" {

int dfm_fm_fwd(const float* e, int64_t batch, int n_fields, int dim, float* out, void* stream) {
    DFM_REQUIRE(e && out, DFM_ERR_INVALID, "dfm_fm_fwd: null argument");
    DFM_REQUIRE(batch >= 0 && n_fields > 0 && dim > 0, DFM_ERR_INVALID, "dfm_fm_fwd: bad shape");
    if (batch == 0) return DFM_OK;
    fm_fwd_kernel<<<(unsigned)ceil_div(batch, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(e, batch, n_fields, dim, out);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_fm_bwd(const float* e, const float* g_out, int64_t batch, int n_fields, int dim, float* g_e, void* stream) {
    DFM_REQUIRE(e && g_out && g_e, DFM_ERR_INVALID, "dfm_fm_bwd: null argument");
    DFM_REQUIRE(batch >= 0 && n_fields > 0 && dim > 0, DFM_ERR_INVALID, "dfm_fm_bwd: bad shape");
    if (batch == 0) return DFM_OK;
    fm_bwd_kernel<<<(unsigned)ceil_div(batch, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(e, g_out, batch, n_fields, dim, g_e);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_sumsq(int n_tensors, const float* const* ptrs, const int64_t* numel, float scale, float* out,
              float* workspace, void* stream);

static int sumsq_partials(int n_tensors, const float* const* ptrs, const int64_t* numel, float* workspace, cudaStream_t st) {
    SumsqArgs a;
    int done = 0, launches = 0;
    while (done < n_tensors || launches == 0) {
        a.nt = 0;
        while (done < n_tensors && a.nt < SS_MAXT) {
            DFM_REQUIRE(numel[done] >= 0 && (numel[done] == 0 || ptrs[done]), DFM_ERR_INVALID, "dfm_sumsq: tensor %d invalid", done);
            a.p[a.nt] = ptrs[done]; a.n[a.nt] = numel[done]; ++a.nt; ++done;
        }
        sumsq_partial_kernel<<<SS_BLOCKS, 256, 0, st>>>(a, workspace, launches > 0);
        ++launches;
    }
    return DFM_OK;
}

int dfm_sumsq_acc(int n_tensors, const float* const* ptrs, const int64_t* numel, double* acc,
                  float* workspace, void* stream) {
    DFM_REQUIRE(n_tensors >= 0 && acc && workspace && (n_tensors == 0 || (ptrs && numel)), DFM_ERR_INVALID,
                "dfm_sumsq_acc: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = sumsq_partials(n_tensors, ptrs, numel, workspace, st);
    if (rc) return rc;
    sumsq_final_f64_kernel<<<1, 256, 0, st>>>(workspace, SS_BLOCKS, acc);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_sumsq(int n_tensors, const float* const* ptrs, const int64_t* numel, float scale, float* out,
              float* workspace, void* stream) {
    DFM_REQUIRE(n_tensors >= 0 && out && workspace && (n_tensors == 0 || (ptrs && numel)), DFM_ERR_INVALID,
                "dfm_sumsq: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = sumsq_partials(n_tensors, ptrs, numel, workspace, st);
    if (rc) return rc;
    sumsq_final_kernel<<<1, 256, 0, st>>>(workspace, SS_BLOCKS, scale, out);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_l2_combine(const double* acc_a, const double* acc_b, float lam, float* out, void* stream) {
    DFM_REQUIRE(acc_a && out, DFM_ERR_INVALID, "dfm_l2_combine: null argument");
    l2_combine_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(acc_a, acc_b, lam, out);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_axpy(const float* p, int64_t numel, float coef, const float* scale_dev, float* g, int accumulate, void* stream) {
    DFM_REQUIRE(p && g && numel >= 0, DFM_ERR_INVALID, "dfm_axpy: null argument");
    if (numel == 0) return DFM_OK;
    long long blocks = ceil_div(numel, 256);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    axpy_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, numel, coef, scale_dev, g, accumulate);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
