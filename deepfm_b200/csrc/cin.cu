// CIN (Compressed Interaction Network) forward / backward, fp32 CUDA-core path.
//
// Reference: deepfm/models/layers/cin.py:66-105 -- per layer
//     outer = einsum("bhd,bfd->bhfd", hidden, x0).reshape(B, H*F, D)     (channel k = h*F + f)
//     act   = relu(conv1d_k1(outer))   (weight (L, K, 1), bias (L))
//     direct, hidden' = act.split([direct, next]) (split_half and not last) else act, act
//     out_part = direct.sum(dim=2)
// The outer product is never materialised.  Every product of the layer is the same GEMM with an
// on-the-fly outer-product operand ("opgemm"):   C[(b,d)][n] = sum_{p,q} U[b,p,d] V[b,q,d] W(p,q,n)
//     forward        U = hidden, V = x0,     n = l,  W(p,q,n) = w[n, p*F + q]
//     d/d hidden     U = g_pre,  V = x0,     n = h,  W(p,q,n) = w[p, n*F + q]
//     d/d x0         U = g_pre,  V = hidden, n = f,  W(p,q,n) = w[p, q*F + n]
// and the weight gradient is its transpose-side twin (dW[l][h*F+f] = sum_(b,d) g_pre * hidden * x0),
// reduced over batch slices in a fixed order (deterministic, no float atomics).
// The tcgen05 tensor-core path (cin_tc.cu) uses the same decomposition.
#include "common.cuh"

namespace dfm {

constexpr int TM = 64;   // rows (b,d) per block tile
constexpr int TN = 64;   // output columns per tile
constexpr int WPAD = 68; // padded row length of the staged W tile (keeps 16-byte alignment)

struct OpGemmArgs {
    const float* U; long long u_bs; int P;
    const float* Vt; long long v_bs; int Q;
    const float* W; long long sp, sq, sn; int N;
    int D; long long M;
    float* out; long long o_bs;
    const float* bias; int relu; int accumulate;
};

__global__ void __launch_bounds__(256)
opgemm_kernel(const __grid_constant__ OpGemmArgs a) {
    extern __shared__ float sm[];
    float* U_s = sm;                         // [P][TM]
    float* V_s = U_s + (size_t)a.P * TM;     // [Q][TM]
    float* W_s = V_s + (size_t)a.Q * TM;     // [Q][WPAD]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * TM;
    const int D = a.D;
    for (int idx = tid; idx < a.P * TM; idx += 256) {
        const int p = idx / TM, rr = idx - p * TM;
        const long long r = m0 + rr;
        float v = 0.f;
        if (r < a.M) { const long long b = r / D; const int d = (int)(r - b * D); v = __ldg(a.U + b * a.u_bs + (long long)p * D + d); }
        U_s[idx] = v;
    }
    for (int idx = tid; idx < a.Q * TM; idx += 256) {
        const int q = idx / TM, rr = idx - q * TM;
        const long long r = m0 + rr;
        float v = 0.f;
        if (r < a.M) { const long long b = r / D; const int d = (int)(r - b * D); v = __ldg(a.Vt + b * a.v_bs + (long long)q * D + d); }
        V_s[idx] = v;
    }
    const bool n_fast = a.sn == 1;
    for (int n0 = 0; n0 < a.N; n0 += TN) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) acc[i][jn] = 0.f;
        for (int p = 0; p < a.P; ++p) {
            __syncthreads();
            for (int idx = tid; idx < a.Q * TN; idx += 256) {
                int q, n;
                if (n_fast) { q = idx / TN; n = idx - q * TN; } else { n = idx / a.Q; q = idx - n * a.Q; }
                float w = 0.f;
                if (n0 + n < a.N) w = __ldg(a.W + p * a.sp + q * a.sq + (long long)(n0 + n) * a.sn);
                W_s[q * WPAD + n] = w;
            }
            __syncthreads();
            const float4 u4 = *reinterpret_cast<const float4*>(U_s + (size_t)p * TM + ty * 4);
            const float u[4] = {u4.x, u4.y, u4.z, u4.w};
            for (int q = 0; q < a.Q; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(V_s + (size_t)q * TM + ty * 4);
                const float4 w4 = *reinterpret_cast<const float4*>(W_s + q * WPAD + tx * 4);
                const float z[4] = {u[0] * v4.x, u[1] * v4.y, u[2] * v4.z, u[3] * v4.w};
                const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(z[i], w[jn], acc[i][jn]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = m0 + ty * 4 + i;
            if (r >= a.M) continue;
            const long long b = r / D;
            const int d = (int)(r - b * D);
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) {
                const int n = n0 + tx * 4 + jn;
                if (n >= a.N) continue;
                float v = acc[i][jn];
                if (a.bias) v += __ldg(a.bias + n);
                if (a.relu) v = fmaxf(v, 0.f);
                float* o = a.out + b * a.o_bs + (long long)n * D + d;
                *o = a.accumulate ? *o + v : v;
            }
        }
    }
}

// out[b, col_off + l] = sum_d act[b, l, d]   for l < direct
__global__ void cin_pool_kernel(const float* __restrict__ act, long long B, int L, int D, int direct,
                                float* __restrict__ out, int out_dim, int col_off) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * direct) return;
    const long long b = i / direct;
    const int l = (int)(i - b * direct);
    const float* p = act + (b * L + l) * D;
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += __ldg(p + d);
    out[b * out_dim + col_off + l] = s;
}

// g_pre = (act > 0) * ( [l < direct] g_out[b, col_off + l] + [l in next range] g_hnext[b, l - next_off, d] )
__global__ void cin_gpre_kernel(const float* __restrict__ act, const float* __restrict__ g_out,
                                const float* __restrict__ g_hnext, long long B, int L, int D, int direct,
                                int out_dim, int col_off, int next_off, int next_n,
                                float* __restrict__ g_pre) {
    const long long n = B * L * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const long long t = i / D;
        const int l = (int)(t % L);
        const long long b = t / L;
        float g = 0.f;
        if (__ldg(act + i) > 0.f) {
            if (l < direct) g = __ldg(g_out + b * out_dim + col_off + l);
            if (g_hnext && l >= next_off && l < next_off + next_n)
                g += __ldg(g_hnext + (b * next_n + (l - next_off)) * D + d);
        }
        g_pre[i] = g;
    }
}

// dW[l][k] (and db[l]) partial sums over one slice of rows (b,d).
struct DwArgs {
    const float* gp;            // (B, L, D)
    const float* hid; long long h_bs; int H;
    const float* x0; int F;
    int L, K, D;
    long long M, slice_rows;
    float* part_w;              // (n_slices, L, K)
    float* part_b;              // (n_slices, L)
};
constexpr int RC = 32;          // rows per reduction chunk

__global__ void __launch_bounds__(256)
cin_dw_kernel(const __grid_constant__ DwArgs a) {
    extern __shared__ float sm[];
    const int F = a.F, D = a.D;
    const int k0 = blockIdx.x * TN, l0 = blockIdx.y * TM;
    const int h_base = k0 / F;
    int k_last = k0 + TN - 1;
    if (k_last >= a.K) k_last = a.K - 1;
    const int nh = k_last / F - h_base + 1;
    float* gp_s = sm;                       // [RC][WPAD]
    float* hid_s = gp_s + RC * WPAD;        // [RC][nh]
    float* x0_s = hid_s + RC * nh;          // [RC][F]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int hj[4], fj[4];
    bool okj[4];
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
        const int k = k0 + tx * 4 + jn;
        okj[jn] = k < a.K;
        const int kk = okj[jn] ? k : k0;
        hj[jn] = kk / F - h_base;
        fj[jn] = kk % F;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) acc[i][jn] = 0.f;
    float accb = 0.f;
    const long long r_lo = (long long)blockIdx.z * a.slice_rows;
    const long long r_hi = (r_lo + a.slice_rows < a.M) ? r_lo + a.slice_rows : a.M;
    for (long long rc = r_lo; rc < r_hi; rc += RC) {
        __syncthreads();
        for (int idx = tid; idx < TM * RC; idx += 256) {
            const int l = idx / RC, rr = idx - l * RC;
            const long long r = rc + rr;
            float v = 0.f;
            if (r < r_hi && l0 + l < a.L) { const long long b = r / D; const int d = (int)(r - b * D); v = __ldg(a.gp + (b * a.L + l0 + l) * D + d); }
            gp_s[rr * WPAD + l] = v;
        }
        for (int idx = tid; idx < nh * RC; idx += 256) {
            const int h = idx / RC, rr = idx - h * RC;
            const long long r = rc + rr;
            float v = 0.f;
            if (r < r_hi) { const long long b = r / D; const int d = (int)(r - b * D); v = __ldg(a.hid + b * a.h_bs + (long long)(h_base + h) * D + d); }
            hid_s[rr * nh + h] = v;
        }
        for (int idx = tid; idx < F * RC; idx += 256) {
            const int f = idx / RC, rr = idx - f * RC;
            const long long r = rc + rr;
            float v = 0.f;
            if (r < r_hi) { const long long b = r / D; const int d = (int)(r - b * D); v = __ldg(a.x0 + (b * F + f) * D + d); }
            x0_s[rr * F + f] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < RC; ++rr) {
            const float4 g4 = *reinterpret_cast<const float4*>(gp_s + rr * WPAD + ty * 4);
            const float g[4] = {g4.x, g4.y, g4.z, g4.w};
            float z[4];
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) z[jn] = hid_s[rr * nh + hj[jn]] * x0_s[rr * F + fj[jn]];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(g[i], z[jn], acc[i][jn]);
        }
        if (blockIdx.x == 0 && tid < TM) {
            for (int rr = 0; rr < RC; ++rr) accb += gp_s[rr * WPAD + tid];
        }
    }
    float* pw = a.part_w + (size_t)blockIdx.z * a.L * a.K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int l = l0 + ty * 4 + i;
        if (l >= a.L) continue;
#pragma unroll
        for (int jn = 0; jn < 4; ++jn)
            if (okj[jn]) pw[(size_t)l * a.K + k0 + tx * 4 + jn] = acc[i][jn];
    }
    if (blockIdx.x == 0 && tid < TM && l0 + tid < a.L) a.part_b[(size_t)blockIdx.z * a.L + l0 + tid] = accb;
}

__global__ void cin_reduce_kernel(const float* __restrict__ part, int n_slices, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int s = 0; s < n_slices; ++s) acc += part[(size_t)s * n + i];
    out[i] = acc;
}

struct CinPlan {
    int n, F, D, split;
    int L[16], direct[16], next[16], H[16], K[16], col_off[16];
    long long act_off[16];   // floats, per sample
    int out_dim;
    long long act_per_sample;
    int Lmax, Hmax;
    long long LKmax;
};

static int cin_plan(int F, int D, int n_layers, const int32_t* sizes, int split_half, CinPlan& c) {
    if (n_layers <= 0 || n_layers > 16 || F <= 0 || D <= 0 || !sizes) { set_error("cin: need 1..16 layers, F, D > 0"); return DFM_ERR_INVALID; }
    c.n = n_layers; c.F = F; c.D = D; c.split = split_half;
    int prev = F, col = 0;
    long long off = 0;
    c.Lmax = 0; c.Hmax = F; c.LKmax = 0;
    for (int i = 0; i < n_layers; ++i) {
        const int L = sizes[i];
        if (L <= 0) { set_error("cin: layer size must be positive"); return DFM_ERR_INVALID; }
        c.L[i] = L; c.H[i] = prev; c.K[i] = prev * F; c.col_off[i] = col; c.act_off[i] = off;
        if (split_half && i < n_layers - 1) { c.direct[i] = L / 2; c.next[i] = L - L / 2; }
        else { c.direct[i] = L; c.next[i] = L; }
        if (c.next[i] <= 0 && i < n_layers - 1) { set_error("cin: layer %d feeds nothing forward", i); return DFM_ERR_INVALID; }
        col += c.direct[i]; off += (long long)L * D; prev = c.next[i];
        if (L > c.Lmax) c.Lmax = L;
        if (c.H[i] > c.Hmax) c.Hmax = c.H[i];
        if ((long long)L * c.K[i] > c.LKmax) c.LKmax = (long long)L * c.K[i];
    }
    c.out_dim = col; c.act_per_sample = off;
    return DFM_OK;
}

static int dw_slices(long long M) {
    long long s = ceil_div(M, 4096);
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    return (int)s;
}

static int launch_opgemm(const OpGemmArgs& a, cudaStream_t st) {
    const size_t smem = ((size_t)(a.P + a.Q) * TM + (size_t)a.Q * WPAD) * 4;
    DFM_REQUIRE(smem <= 200 * 1024, DFM_ERR_UNSUPPORTED, "cin: %d + %d channels need %zu B shared memory", a.P, a.Q, smem);
    if (smem > 48 * 1024) DFM_CHECK_CUDA(cudaFuncSetAttribute(opgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    opgemm_kernel<<<(unsigned)ceil_div(a.M, TM), 256, smem, st>>>(a);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

// tcgen05 path (cin_tc.cu)
int cin_layer_fwd_tc(const float* x0, long long x_bs, const float* hid, long long h_bs, const float* w,
                     const float* bias, float* act, long long B, int F, int H, int D, int L, float* wpad,
                     cudaStream_t st);
size_t cin_tc_wpad_floats(int F, int Hmax, int Lmax);
bool cin_tc_bwd_supported(long long B, int F, int H, int D, int L);
int cin_gpre_fused(const float* act, const float* g_out, const float* g_hnext, long long B, int L, int D, int direct,
                   int out_dim, int col_off, int next_off, int next_n, float* bwd_scratch, float* dw_scratch, cudaStream_t st);
size_t cin_tc_bwd_scratch_floats(long long B, int F, int D, int Hmax, int Lmax);
size_t cin_tc_dw_scratch_floats(long long B, int F, int D, int Hmax, int Lmax);
int cin_layer_dw_tc(const float* g_pre, const float* x0, long long x_bs, const float* hid, long long h_bs, float* gw,
                    float* gb, long long B, int F, int H, int D, int L, float* scratch, cudaStream_t st);
int cin_layer_bwd_data_tc(const float* g_pre, const float* x0, long long x_bs, const float* hid, long long h_bs,
                          const float* w, float* g_hid, long long gh_bs, int gh_accumulate, float* g_x0, long long B,
                          int F, int H, int D, int L, float* scratch, cudaStream_t st);

}  // namespace dfm

using namespace dfm;

extern "C" {

int dfm_cin_sizes(int n_fields, int dim, int n_layers, const int32_t* layer_sizes, int split_half, int64_t batch,
                  int64_t out[4]) {
    DFM_REQUIRE(out, DFM_ERR_INVALID, "dfm_cin_sizes: out is null");
    CinPlan c;
    int rc = cin_plan(n_fields, dim, n_layers, layer_sizes, split_half, c);
    if (rc) return rc;
    const long long M = batch * dim;
    out[0] = c.out_dim;
    out[1] = (c.act_per_sample * batch + (long long)cin_tc_wpad_floats(c.F, c.Hmax, c.Lmax)) * 4;   // activations + padded-W scratch
    size_t ws = 0;
    ws += align_up((size_t)batch * c.Lmax * dim * 4, 256);      // g_pre
    ws += 2 * align_up((size_t)batch * c.Hmax * dim * 4, 256);  // g_hidden ping-pong
    ws += align_up((size_t)dw_slices(M) * c.LKmax * 4, 256);    // dW partials
    ws += align_up((size_t)dw_slices(M) * c.Lmax * 4, 256);     // db partials
    ws += align_up(cin_tc_bwd_scratch_floats(batch, c.F, dim, c.Hmax, c.Lmax) * 4, 256);   // tcgen05 backward scratch
    ws += align_up(cin_tc_dw_scratch_floats(batch, c.F, dim, c.Hmax, c.Lmax) * 4, 256);
    out[2] = (int64_t)ws;                                       // bytes of the backward workspace
    out[3] = c.act_per_sample;
    return DFM_OK;
}

int dfm_cin_fwd(const float* x0, int64_t batch, int n_fields, int dim, int n_layers,
                const int32_t* layer_sizes, int split_half, const float* const* weights,
                const float* const* biases, int precision, float* out, float* acts, void* stream) {
    DFM_REQUIRE(weights && biases && layer_sizes, DFM_ERR_INVALID, "dfm_cin_fwd: null argument");
    DFM_REQUIRE(batch >= 0, DFM_ERR_INVALID, "dfm_cin_fwd: negative batch");
    DFM_REQUIRE(precision == 0 || precision == 1, DFM_ERR_UNSUPPORTED, "dfm_cin_fwd: precision %d not built", precision);
    CinPlan c;
    int rc = cin_plan(n_fields, dim, n_layers, layer_sizes, split_half, c);
    if (rc) return rc;
    if (batch == 0) return DFM_OK;
    DFM_REQUIRE(x0 && out && acts, DFM_ERR_INVALID, "dfm_cin_fwd: null tensor");
    float* wpad = acts + c.act_per_sample * batch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int F = c.F, D = c.D;
    const float* hid = x0;
    long long h_bs = (long long)F * D;
    for (int i = 0; i < c.n; ++i) {
        float* act = acts + c.act_off[i] * batch;
        OpGemmArgs a;
        a.U = hid; a.u_bs = h_bs; a.P = c.H[i];
        a.Vt = x0; a.v_bs = (long long)F * D; a.Q = F;
        a.W = weights[i]; a.sp = F; a.sq = 1; a.sn = c.K[i]; a.N = c.L[i];
        a.D = D; a.M = batch * D;
        a.out = act; a.o_bs = (long long)c.L[i] * D;
        a.bias = biases[i]; a.relu = 1; a.accumulate = 0;
        DFM_REQUIRE(weights[i] && biases[i], DFM_ERR_INVALID, "dfm_cin_fwd: layer %d weight/bias null", i);
        if (precision == 1)   // tensor cores: TF32 inputs, FP32 accumulate
            rc = cin_layer_fwd_tc(x0, (long long)F * D, hid, h_bs, weights[i], biases[i], act, batch, F, c.H[i], D, c.L[i], wpad, st);
        else
            rc = launch_opgemm(a, st);
        if (rc) return rc;
        const long long np = batch * c.direct[i];
        cin_pool_kernel<<<(unsigned)ceil_div(np, 256), 256, 0, st>>>(act, batch, c.L[i], D, c.direct[i], out, c.out_dim, c.col_off[i]);
        DFM_CHECK_LAUNCH();
        // next hidden = the "next" channels of this layer's activation
        const int next_off = (c.split && i < c.n - 1) ? c.direct[i] : 0;
        hid = act + (long long)next_off * D;
        h_bs = (long long)c.L[i] * D;
    }
    return DFM_OK;
}

int dfm_cin_bwd(const float* x0, const float* g_out, int64_t batch, int n_fields, int dim, int n_layers,
                const int32_t* layer_sizes, int split_half, const float* const* weights, int precision,
                const float* acts, float* g_x0, float* const* g_weights, float* const* g_biases,
                void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(weights && g_weights && g_biases && layer_sizes, DFM_ERR_INVALID, "dfm_cin_bwd: null argument");
    DFM_REQUIRE(batch >= 0, DFM_ERR_INVALID, "dfm_cin_bwd: negative batch");
    DFM_REQUIRE(precision == 0 || precision == 1, DFM_ERR_UNSUPPORTED, "dfm_cin_bwd: precision %d not built", precision);
    CinPlan c;
    int rc = cin_plan(n_fields, dim, n_layers, layer_sizes, split_half, c);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int F = c.F, D = c.D;
    if (batch == 0) {
        for (int i = 0; i < c.n; ++i) {
            DFM_CHECK_CUDA(cudaMemsetAsync(g_weights[i], 0, (size_t)c.L[i] * c.K[i] * 4, st));
            DFM_CHECK_CUDA(cudaMemsetAsync(g_biases[i], 0, (size_t)c.L[i] * 4, st));
        }
        return DFM_OK;
    }
    DFM_REQUIRE(x0 && g_out && acts && g_x0 && workspace, DFM_ERR_INVALID, "dfm_cin_bwd: null tensor");
    const long long M = batch * D;
    const int n_slices = dw_slices(M);
    char* ws = static_cast<char*>(workspace);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return ws + o; };
    float* g_pre = reinterpret_cast<float*>(take((size_t)batch * c.Lmax * D * 4));
    float* g_hid[2];
    g_hid[0] = reinterpret_cast<float*>(take((size_t)batch * c.Hmax * D * 4));
    g_hid[1] = reinterpret_cast<float*>(take((size_t)batch * c.Hmax * D * 4));
    float* part_w = reinterpret_cast<float*>(take((size_t)n_slices * c.LKmax * 4));
    float* part_b = reinterpret_cast<float*>(take((size_t)n_slices * c.Lmax * 4));
    float* tc_scratch = reinterpret_cast<float*>(take(cin_tc_bwd_scratch_floats(batch, c.F, D, c.Hmax, c.Lmax) * 4));
    float* dw_scratch = reinterpret_cast<float*>(take(cin_tc_dw_scratch_floats(batch, c.F, D, c.Hmax, c.Lmax) * 4));
    DFM_REQUIRE(off <= workspace_bytes, DFM_ERR_WORKSPACE, "dfm_cin_bwd: workspace %zu < %zu", workspace_bytes, off);
    DFM_CHECK_CUDA(cudaMemsetAsync(g_x0, 0, (size_t)batch * F * D * 4, st));
    const float* g_hnext = nullptr;   // gradient w.r.t. the hidden input of layer i+1
    for (int i = c.n - 1; i >= 0; --i) {
        const float* act = acts + c.act_off[i] * batch;
        const int L = c.L[i], H = c.H[i], K = c.K[i];
        const int next_off = (c.split && i < c.n - 1) ? c.direct[i] : 0;
        const int next_n = (i < c.n - 1) ? c.next[i] : 0;
        const long long tot = batch * L * D;
        long long gb = ceil_div(tot, 256);
        if (gb > 16LL * sm_count()) gb = 16LL * sm_count();
        // tensor-core path: g_pre is produced once, directly in the two layouts its GEMMs read
        const bool fused = precision == 1 && cin_tc_bwd_supported(batch, F, H, D, L);
        if (fused) {
            rc = cin_gpre_fused(act, g_out, g_hnext, batch, L, D, c.direct[i], c.out_dim, c.col_off[i], next_off, next_n,
                                tc_scratch, dw_scratch, st);
            if (rc) return rc;
        } else {
            cin_gpre_kernel<<<(unsigned)gb, 256, 0, st>>>(act, g_out, g_hnext, batch, L, D, c.direct[i], c.out_dim,
                                                          c.col_off[i], next_off, next_n, g_pre);
            DFM_CHECK_LAUNCH();
        }
        const float* g_pre_in = fused ? nullptr : g_pre;
        // hidden input of this layer
        const float* hid = x0;
        long long h_bs = (long long)F * D;
        if (i > 0) {
            const int poff = (c.split) ? c.direct[i - 1] : 0;
            hid = acts + c.act_off[i - 1] * batch + (long long)poff * D;
            h_bs = (long long)c.L[i - 1] * D;
        }
        // weight / bias gradients
        bool dw_done = false;
        if (precision == 1) {
            rc = cin_layer_dw_tc(g_pre_in, x0, (long long)F * D, hid, h_bs, g_weights[i], g_biases[i], batch, F, H, D, L, dw_scratch, st);
            if (rc == DFM_OK) dw_done = true;
            else if (rc != DFM_ERR_UNSUPPORTED || fused) return rc;
        }
        if (!dw_done) {
        DwArgs d;
        d.gp = g_pre; d.hid = hid; d.h_bs = h_bs; d.H = H; d.x0 = x0; d.F = F; d.L = L; d.K = K; d.D = D;
        d.M = M; d.slice_rows = ceil_div(ceil_div(M, n_slices), RC) * RC;
        d.part_w = part_w; d.part_b = part_b;
        const int real_slices = (int)ceil_div(M, d.slice_rows);
        const int nh_max = (TN - 1) / F + 2;
        const size_t smem = ((size_t)RC * WPAD + (size_t)RC * nh_max + (size_t)RC * F) * 4;
        DFM_REQUIRE(smem <= 200 * 1024, DFM_ERR_UNSUPPORTED, "cin dW: %zu B shared memory", smem);
        if (smem > 48 * 1024) DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cin_dw_kernel<<<dim3((unsigned)ceil_div(K, TN), (unsigned)ceil_div(L, TM), real_slices), 256, smem, st>>>(d);
        cin_reduce_kernel<<<(unsigned)ceil_div((long long)L * K, 256), 256, 0, st>>>(part_w, real_slices, (long long)L * K, g_weights[i]);
        cin_reduce_kernel<<<(unsigned)ceil_div(L, 256), 256, 0, st>>>(part_b, real_slices, L, g_biases[i]);
        DFM_CHECK_LAUNCH();
        }
        float* gh = g_hid[i & 1];
        if (precision == 1) {   // tensor cores: gz = g_pre W stays in TMEM, contracted per row in the epilogue
            rc = cin_layer_bwd_data_tc(g_pre_in, x0, (long long)F * D, hid, h_bs, weights[i], i == 0 ? g_x0 : gh,
                                       i == 0 ? (long long)F * D : (long long)H * D, i == 0 ? 1 : 0, g_x0, batch, F, H, D, L,
                                       tc_scratch, st);
            if (rc == DFM_OK) { g_hnext = gh; continue; }
            if (rc != DFM_ERR_UNSUPPORTED || fused) return rc;   // unsupported tile shape: fp32 CUDA-core path below
        }
        // d/d hidden  (layer 0: the hidden input is x0 itself -> accumulate into g_x0)
        OpGemmArgs a;
        a.U = g_pre; a.u_bs = (long long)L * D; a.P = L;
        a.Vt = x0; a.v_bs = (long long)F * D; a.Q = F;
        a.W = weights[i]; a.sp = K; a.sq = 1; a.sn = F; a.N = H;
        a.D = D; a.M = M; a.bias = nullptr; a.relu = 0;
        if (i == 0) { a.out = g_x0; a.o_bs = (long long)F * D; a.accumulate = 1; }
        else { a.out = gh; a.o_bs = (long long)H * D; a.accumulate = 0; }
        rc = launch_opgemm(a, st);
        if (rc) return rc;
        // d/d x0
        a.U = g_pre; a.u_bs = (long long)L * D; a.P = L;
        a.Vt = hid; a.v_bs = h_bs; a.Q = H;
        a.sp = K; a.sq = F; a.sn = 1; a.N = F;
        a.out = g_x0; a.o_bs = (long long)F * D; a.accumulate = 1;
        rc = launch_opgemm(a, st);
        if (rc) return rc;
        g_hnext = gh;
    }
    return DFM_OK;
}

}  // extern "C"
