// Field self-attention block, warp-per-sample kernels for the shape family of BASELINE config 3
// (AttentionDeepFM on the ML-100K schema: embed_dim D = 16, head dim A / heads = 16, F <= 16 fields).
//
// Reference: deepfm/models/layers/attention.py:91-120 (_AttentionBlock.forward) and its autograd.
//
// One warp owns one sample at a time and never meets a block barrier.  Lane l holds row r = l / 2 (a field)
// and the 8 columns [8 hf, 8 hf + 8), hf = l % 2, of every 16 x 16 tile (x, Q_h, K_h, V_h, P_h, O_h, y and
// their gradients): a row is completed with one shuffle between the two lanes of a pair, tiles another row
// needs (K, V forward; Q, P, gS, gO backward) sit in a 7.5 KB per-warp shared-memory scratch, the weights of
// the block are staged once per CTA (transposed for x W^T products, natural for g W products).  The heads are
// processed one after the other and the output projection is accumulated head by head, so nothing wider
// than 16 x 16 ever exists.
// Backward: pass A recomputes y for the LayerNorm backward, pass B recomputes Q/K/V/P per head and
// back-propagates; the per-row gradients [gQ | gK | gV], O and g_r go to a global scratch once and the
// parameter gradients are ONE tall-skinny GEMM each over those rows (tsgemm_kernel: per-slice partial sums,
// added in slice order -- deterministic, no float atomics).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace dfm {

constexpr int A16_ST = 20;   // row stride (floats) of the per-warp 16 x 16 tiles (80 B: rows spread over the banks)
constexpr int A16_WARPS = 8;

struct Attn16Args {
    const float* x;
    const float* g_out;       // backward only
    float* out;               // forward: y ; backward: g_x
    long long B;
    int F, A, H, residual;
    float eps, inv_scale;
    const float *Wq, *bq, *Wk, *bk, *Wv, *bv, *Wo, *bo, *gamma, *beta;
    float* G;                 // backward: (B*F, 3A)  [gQ | gK | gV]
    float* O;                 // backward: (B*F, A)   merged heads (input of W_out)
    float* GR;                // backward: (B*F, 16)  gradient w.r.t. the W_out output (after the LayerNorm backward)
    float* wpart;             // backward: (n_warps, 32) per-warp d gamma | d beta
};

struct Attn16W {              // block-shared weights in shared memory
    const float *WqT, *WkT, *WvT;   // [d][a]   (x W^T: fixed d, consecutive a)
    const float* WoT;               // [a][d]
    const float *bq, *bk, *bv, *bo, *gamma, *beta;
    const float *WqN, *WkN, *WvN;   // backward: [a][d]   (g W: fixed a, consecutive d)
    const float* WoN;               // backward: [d][a]
};

__host__ __device__ inline size_t attn16_weight_floats(int A, bool bwd) {
    return (size_t)4 * 16 * A + 3 * A + 48 + (bwd ? (size_t)4 * 16 * A : 0);
}
__host__ __device__ inline size_t attn16_smem_bytes(int A, bool bwd) {
    return (attn16_weight_floats(A, bwd) + (size_t)A16_WARPS * (bwd ? 6 : 2) * 16 * A16_ST) * 4;
}

__device__ __forceinline__ float* attn16_stage_weights(float* sm, const Attn16Args& a, bool bwd, Attn16W& w) {
    const int A = a.A, nt = blockDim.x, tid = threadIdx.x;
    float* WqT = sm; float* WkT = WqT + 16 * A; float* WvT = WkT + 16 * A; float* WoT = WvT + 16 * A;
    float* bq = WoT + 16 * A; float* bk = bq + A; float* bv = bk + A; float* bo = bv + A;
    float* gamma = bo + 16; float* beta = gamma + 16;
    float* p = beta + 16;
    float* WqN = p; float* WkN = WqN + 16 * A; float* WvN = WkN + 16 * A; float* WoN = WvN + 16 * A;
    for (int i = tid; i < A * 16; i += nt) {
        const int aa = i >> 4, d = i & 15;                 // Wq[a][d]
        const float q = __ldg(a.Wq + i), k = __ldg(a.Wk + i), v = __ldg(a.Wv + i);
        WqT[d * A + aa] = q; WkT[d * A + aa] = k; WvT[d * A + aa] = v;
        if (bwd) { WqN[i] = q; WkN[i] = k; WvN[i] = v; }
        const int d2 = i / A, a2 = i - d2 * A;             // Wo[d][a]
        const float o = __ldg(a.Wo + i);
        WoT[a2 * 16 + d2] = o;
        if (bwd) WoN[i] = o;
    }
    for (int i = tid; i < A; i += nt) { bq[i] = __ldg(a.bq + i); bk[i] = __ldg(a.bk + i); bv[i] = __ldg(a.bv + i); }
    for (int i = tid; i < 16; i += nt) {
        bo[i] = __ldg(a.bo + i);
        gamma[i] = a.residual ? __ldg(a.gamma + i) : 1.f;
        beta[i] = a.residual ? __ldg(a.beta + i) : 0.f;
    }
    w.WqT = WqT; w.WkT = WkT; w.WvT = WvT; w.WoT = WoT; w.bq = bq; w.bk = bk; w.bv = bv; w.bo = bo;
    w.gamma = gamma; w.beta = beta; w.WqN = WqN; w.WkN = WkN; w.WvN = WvN; w.WoN = WoN;
    return bwd ? WoN + 16 * A : p;
}

// the full 16-wide row from the two 8-wide halves of a lane pair (compile-time register indices only)
__device__ __forceinline__ void full16(const float (&own)[8], int hf, float (&full)[16]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float other = __shfl_xor_sync(0xffffffffu, own[i], 1);
        full[i] = hf ? other : own[i];
        full[8 + i] = hf ? own[i] : other;
    }
}
__device__ __forceinline__ float pair_sum(float v) { return v + __shfl_xor_sync(0xffffffffu, v, 1); }
// sum over the 16 rows (lanes of equal hf)
__device__ __forceinline__ float rows_sum(float v) {
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// acc[0..8) += s * row[0..8)
__device__ __forceinline__ void axpy8(float s, const float* row, float (&acc)[8]) {
    float w[8];
    ld8(row, w);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(s, w[i], acc[i]);
}
__device__ __forceinline__ float dot16(const float (&f)[16], const float* row) {
    float a[8], b[8];
    ld8(row, a); ld8(row + 8, b);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(f[i], a[i], s);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(f[8 + i], b[i], s);
    return s;
}

// One head of the forward for the warp's sample: q, p (own half-rows), o = P V (own half-row); K, V (and, for the
// backward, Q and P) are left in the per-warp tiles.  Ends with the tiles readable by every lane.
template <bool BWD>
__device__ __forceinline__ void attn16_head(const Attn16W& w, int A, int h, int F, float inv_scale, const float (&xr)[16],
                                            int r, int hf, float* sq, float* sk, float* sv, float* sp,
                                            float (&q)[8], float (&p)[8], float (&o)[8]) {
    const int col = h * 16 + 8 * hf;
    float k[8], v[8];
    ld8(w.bq + col, q); ld8(w.bk + col, k); ld8(w.bv + col, v);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        axpy8(xr[kk], w.WqT + kk * A + col, q);
        axpy8(xr[kk], w.WkT + kk * A + col, k);
        axpy8(xr[kk], w.WvT + kk * A + col, v);
    }
    __syncwarp();                                        // the previous head's readers are done with the tiles
    st8(sk + r * A16_ST + 8 * hf, k);
    st8(sv + r * A16_ST + 8 * hf, v);
    if (BWD) st8(sq + r * A16_ST + 8 * hf, q);
    __syncwarp();
    float qf[16];
    full16(q, hf, qf);
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const int j = 8 * hf + jj;
        float s = dot16(qf, sk + j * A16_ST) * inv_scale;
        if (j >= F) s = -INFINITY;
        p[jj] = s;
        mx = fmaxf(mx, s);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) { p[jj] = expf(p[jj] - mx); sum += p[jj]; }
    sum = pair_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) p[jj] *= inv;
    if (BWD) st8(sp + r * A16_ST + 8 * hf, p);
    float pf[16];
    full16(p, hf, pf);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) axpy8(pf[j], sv + j * A16_ST + 8 * hf, o);
}

// y[r][8hf..] += O_h[r][:] W_out[:, h*16 ..]^T
__device__ __forceinline__ void attn16_out_proj(const Attn16W& w, int h, int hf, const float (&o)[8], float (&y)[8]) {
    float of[16];
    full16(o, hf, of);
#pragma unroll
    for (int aa = 0; aa < 16; ++aa) axpy8(of[aa], w.WoT + (h * 16 + aa) * 16 + 8 * hf, y);
}

__global__ void __launch_bounds__(A16_WARPS * 32, 2)
attn16_fwd_kernel(const __grid_constant__ Attn16Args a) {
    extern __shared__ __align__(16) float sm16[];
    Attn16W w;
    float* tiles = attn16_stage_weights(sm16, a, false, w);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, r = lane >> 1, hf = lane & 1;
    float* sk = tiles + (size_t)wib * 2 * 16 * A16_ST;
    float* sv = sk + 16 * A16_ST;
    const int F = a.F, A = a.A;
    const long long gw = (long long)blockIdx.x * A16_WARPS + wib, nw = (long long)gridDim.x * A16_WARPS;
    for (long long b = gw; b < a.B; b += nw) {
        float xo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < F) ld8(a.x + ((size_t)b * F + r) * 16 + 8 * hf, xo);
        float xr[16];
        full16(xo, hf, xr);
        float y[8];
        ld8(w.bo + 8 * hf, y);
        for (int h = 0; h < a.H; ++h) {
            float q[8], p[8], o[8];
            attn16_head<false>(w, A, h, F, a.inv_scale, xr, r, hf, nullptr, sk, sv, nullptr, q, p, o);
            attn16_out_proj(w, h, hf, o, y);
        }
        if (a.residual) {   // LayerNorm(y + x): biased variance, eps inside the square root
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { y[i] += xo[i]; s += y[i]; }
            const float mean = pair_sum(s) * (1.f / 16.f);
            float var = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float c = y[i] - mean; var = fmaf(c, c, var); }
            const float rstd = rsqrtf(pair_sum(var) * (1.f / 16.f) + a.eps);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = (y[i] - mean) * rstd * w.gamma[8 * hf + i] + w.beta[8 * hf + i];
        }
        if (r < F) st8(a.out + ((size_t)b * F + r) * 16 + 8 * hf, y);
    }
}

__global__ void __launch_bounds__(A16_WARPS * 32, 2)
attn16_bwd_kernel(const __grid_constant__ Attn16Args a) {
    extern __shared__ __align__(16) float sm16[];
    Attn16W w;
    float* tiles = attn16_stage_weights(sm16, a, true, w);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, r = lane >> 1, hf = lane & 1;
    float* sq = tiles + (size_t)wib * 6 * 16 * A16_ST;
    float* sk = sq + 16 * A16_ST; float* sv = sk + 16 * A16_ST; float* sp = sv + 16 * A16_ST;
    float* sgs = sp + 16 * A16_ST; float* sgo = sgs + 16 * A16_ST;
    const int F = a.F, A = a.A;
    const long long gw = (long long)blockIdx.x * A16_WARPS + wib, nw = (long long)gridDim.x * A16_WARPS;
    float acc_ga[8], acc_be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc_ga[i] = 0.f; acc_be[i] = 0.f; }
    for (long long b = gw; b < a.B; b += nw) {
        const size_t row = (size_t)b * F + r;
        float xo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gy[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < F) { ld8(a.x + row * 16 + 8 * hf, xo); ld8(a.g_out + row * 16 + 8 * hf, gy); }
        float xr[16];
        full16(xo, hf, xr);
        float gr[8];
        if (a.residual) {
            // pass A: y = W_out(heads) + b_out, the LayerNorm input is y + x
            float y[8];
            ld8(w.bo + 8 * hf, y);
            for (int h = 0; h < a.H; ++h) {
                float q[8], p[8], o[8];
                attn16_head<false>(w, A, h, F, a.inv_scale, xr, r, hf, nullptr, sk, sv, nullptr, q, p, o);
                attn16_out_proj(w, h, hf, o, y);
            }
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { y[i] += xo[i]; s += y[i]; }
            const float mean = pair_sum(s) * (1.f / 16.f);
            float var = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float c = y[i] - mean; var = fmaf(c, c, var); }
            const float rstd = rsqrtf(pair_sum(var) * (1.f / 16.f) + a.eps);
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xh = (y[i] - mean) * rstd;
                y[i] = xh;
                acc_ga[i] += rows_sum(gy[i] * xh);              // rows r >= F carry gy = 0
                acc_be[i] += rows_sum(gy[i]);
                const float t = gy[i] * w.gamma[8 * hf + i];
                gr[i] = t;
                m1 += t; m2 = fmaf(t, xh, m2);
            }
            m1 = pair_sum(m1) * (1.f / 16.f); m2 = pair_sum(m2) * (1.f / 16.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) gr[i] = rstd * (gr[i] - m1 - y[i] * m2);
            if (r >= F) {
#pragma unroll
                for (int i = 0; i < 8; ++i) gr[i] = 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) gr[i] = gy[i];
        }
        float gx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) gx[i] = a.residual ? gr[i] : 0.f;
        if (r < F) st8(a.GR + row * 16 + 8 * hf, gr);
        float grf[16];
        full16(gr, hf, grf);
        // pass B: per head, recompute and back-propagate
        for (int h = 0; h < a.H; ++h) {
            const int col = h * 16 + 8 * hf;
            float q[8], p[8], o[8];
            attn16_head<true>(w, A, h, F, a.inv_scale, xr, r, hf, sq, sk, sv, sp, q, p, o);
            if (r < F) st8(a.O + row * A + col, o);
            float go[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // gO_h = g_r W_out[:, head]
#pragma unroll
            for (int d = 0; d < 16; ++d) axpy8(grf[d], w.WoN + d * A + col, go);
            st8(sgo + r * A16_ST + 8 * hf, go);
            float gof[16];
            full16(go, hf, gof);
            float gs[8], dot = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {                              // gP[r][j] = gO_h[r] . V_h[j]
                gs[jj] = dot16(gof, sv + (8 * hf + jj) * A16_ST);
                dot = fmaf(gs[jj], p[jj], dot);
            }
            dot = pair_sum(dot);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) gs[jj] = p[jj] * (gs[jj] - dot) * a.inv_scale;
            st8(sgs + r * A16_ST + 8 * hf, gs);
            __syncwarp();
            float gsf[16];
            full16(gs, hf, gsf);
            float gq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gk[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
                  gv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                axpy8(gsf[j], sk + j * A16_ST + 8 * hf, gq);              // gQ[r] = sum_j gS[r][j] K[j]
                axpy8(sgs[j * A16_ST + r], sq + j * A16_ST + 8 * hf, gk); // gK[r] = sum_i gS[i][r] Q[i]
                axpy8(sp[j * A16_ST + r], sgo + j * A16_ST + 8 * hf, gv); // gV[r] = sum_i P[i][r] gO[i]
            }
            if (r < F) {
                float* g = a.G + row * 3 * A + col;
                st8(g, gq); st8(g + A, gk); st8(g + 2 * A, gv);
            }
            float f16[16];
            full16(gq, hf, f16);
#pragma unroll
            for (int aa = 0; aa < 16; ++aa) axpy8(f16[aa], w.WqN + (h * 16 + aa) * 16 + 8 * hf, gx);
            full16(gk, hf, f16);
#pragma unroll
            for (int aa = 0; aa < 16; ++aa) axpy8(f16[aa], w.WkN + (h * 16 + aa) * 16 + 8 * hf, gx);
            full16(gv, hf, f16);
#pragma unroll
            for (int aa = 0; aa < 16; ++aa) axpy8(f16[aa], w.WvN + (h * 16 + aa) * 16 + 8 * hf, gx);
        }
        if (r < F) st8(a.out + row * 16 + 8 * hf, gx);
    }
    if (lane < 2) {   // lanes 0 / 1: d gamma, d beta of columns [8 hf, 8 hf + 8) summed over this warp's samples
        float* dst = a.wpart + (size_t)gw * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) { dst[8 * hf + i] = acc_ga[i]; dst[16 + 8 * hf + i] = acc_be[i]; }
    }
}

// C[n][m] = sum_rows L[row][n] * R[row][m]  and  csum[n] = sum_rows L[row][n], for a slice of rows per block:
// partials[slice] = [C (N*M) | csum (N)].  N, M multiples of 4; thread tiles of 4 x 4, RC rows per smem chunk.
constexpr int TS_RC = 32;
template <int TS_LV, int TS_RV>   // float4 of L / of R per thread and chunk: N <= 32 * TS_LV, M <= 32 * TS_RV
__global__ void __launch_bounds__(256)
tsgemm_kernel(const float* __restrict__ L, int ldl, int N, const float* __restrict__ R, int ldr, int M,
              long long rows, long long slice_rows, float* __restrict__ partials) {
    extern __shared__ __align__(16) float ts_sm[];
    float* Ls = ts_sm;                    // [TS_RC][N]
    float* Rs = Ls + TS_RC * N;           // [TS_RC][M]
    const int tid = threadIdx.x;
    const int tm = M >> 2, n_tiles = (N >> 2) * tm;
    const int t0 = tid, t1 = tid + 256;   // up to two 4 x 4 tiles per thread
    float acc0[4][4], acc1[4][4], cs0[4] = {0.f, 0.f, 0.f, 0.f}, cs1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc0[i][j] = 0.f; acc1[i][j] = 0.f; }
    const int n0 = (t0 / tm) << 2, m0 = (t0 % tm) << 2, n1 = (t1 / tm) << 2, m1 = (t1 % tm) << 2;
    const bool on0 = t0 < n_tiles, on1 = t1 < n_tiles;
    const long long r_lo = (long long)blockIdx.x * slice_rows;
    const long long r_hi = r_lo + slice_rows < rows ? r_lo + slice_rows : rows;
    // register double buffer: the next chunk's global loads are in flight while the current chunk is multiplied
    const int nl4 = N >> 2, n_l = TS_RC * nl4, n_r = TS_RC * tm;
    float4 lbuf[TS_LV], rbuf[TS_RV];
    auto fetch = [&](long long rc) {
        const int nr = (int)(r_hi - rc < TS_RC ? r_hi - rc : TS_RC);
#pragma unroll
        for (int q = 0; q < TS_LV; ++q) {
            const int i = tid + q * 256;
            lbuf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_l) {
                const int rr = i / nl4, c = (i - rr * nl4) << 2;
                if (rr < nr) lbuf[q] = __ldcs(reinterpret_cast<const float4*>(L + (size_t)(rc + rr) * ldl + c));
            }
        }
#pragma unroll
        for (int q = 0; q < TS_RV; ++q) {
            const int i = tid + q * 256;
            rbuf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_r) {
                const int rr = i / tm, c = (i - rr * tm) << 2;
                if (rr < nr) rbuf[q] = __ldcs(reinterpret_cast<const float4*>(R + (size_t)(rc + rr) * ldr + c));
            }
        }
    };
    if (r_lo < r_hi) fetch(r_lo);
    for (long long rc = r_lo; rc < r_hi; rc += TS_RC) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TS_LV; ++q) {
            const int i = tid + q * 256;
            if (i < n_l) { const int rr = i / nl4, c = (i - rr * nl4) << 2; *reinterpret_cast<float4*>(Ls + rr * N + c) = lbuf[q]; }
        }
#pragma unroll
        for (int q = 0; q < TS_RV; ++q) {
            const int i = tid + q * 256;
            if (i < n_r) { const int rr = i / tm, c = (i - rr * tm) << 2; *reinterpret_cast<float4*>(Rs + rr * M + c) = rbuf[q]; }
        }
        __syncthreads();
        if (rc + TS_RC < r_hi) fetch(rc + TS_RC);
        if (on0) {
#pragma unroll 4
            for (int rr = 0; rr < TS_RC; ++rr) {
                const float4 l4 = *reinterpret_cast<const float4*>(Ls + rr * N + n0);
                const float4 r4 = *reinterpret_cast<const float4*>(Rs + rr * M + m0);
                const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    cs0[i] += lv[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc0[i][j] = fmaf(lv[i], rv[j], acc0[i][j]);
                }
            }
        }
        if (on1) {
#pragma unroll 4
            for (int rr = 0; rr < TS_RC; ++rr) {
                const float4 l4 = *reinterpret_cast<const float4*>(Ls + rr * N + n1);
                const float4 r4 = *reinterpret_cast<const float4*>(Rs + rr * M + m1);
                const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, rv[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    cs1[i] += lv[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc1[i][j] = fmaf(lv[i], rv[j], acc1[i][j]);
                }
            }
        }
    }
    float* out = partials + (size_t)blockIdx.x * ((size_t)N * M + N);
    if (on0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) out[(size_t)(n0 + i) * M + m0 + j] = acc0[i][j];
            if (m0 == 0) out[(size_t)N * M + n0 + i] = cs0[i];
        }
    }
    if (on1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) out[(size_t)(n1 + i) * M + m1 + j] = acc1[i][j];
            if (m1 == 0) out[(size_t)N * M + n1 + i] = cs1[i];
        }
    }
}

// out segments <- sum over partial rows, in row order.  seg: up to 8 (destination, source offset, length) triples.
struct SegOut { float* dst[8]; int off[8]; int len[8]; int n; };
__global__ void seg_reduce_kernel(const float* __restrict__ partials, int n_rows, int row_len,
                                  const __grid_constant__ SegOut so) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < so.n; ++s) {
        if (i >= so.off[s] && i < so.off[s] + so.len[s]) {
            float acc = 0.f;
#pragma unroll 8
            for (int p = 0; p < n_rows; ++p) acc += partials[(size_t)p * row_len + i];
            if (so.dst[s]) so.dst[s][i - so.off[s]] = acc;
            return;
        }
    }
}

bool attn16_supported(int F, int D, int A, int heads) {
    return D == 16 && F >= 1 && F <= 16 && heads >= 1 && A == 16 * heads && A <= 128;
}

static void attn16_fill(Attn16Args& a, const float* x, const float* g_out, float* out, int64_t B, int F, int A, int heads,
                        int residual, const float* const* p) {
    memset(&a, 0, sizeof(a));
    a.x = x; a.g_out = g_out; a.out = out; a.B = B; a.F = F; a.A = A; a.H = heads; a.residual = residual;
    a.eps = 1e-5f; a.inv_scale = 1.f / sqrtf(16.f);
    a.Wq = p[0]; a.bq = p[1]; a.Wk = p[2]; a.bk = p[3]; a.Wv = p[4]; a.bv = p[5]; a.Wo = p[6]; a.bo = p[7];
    a.gamma = residual ? p[8] : nullptr; a.beta = residual ? p[9] : nullptr;
}

static int attn16_grid(int64_t B) {
    long long blocks = ceil_div(B > 0 ? B : 1, A16_WARPS);
    if (blocks > 2LL * sm_count()) blocks = 2LL * sm_count();
    return (int)blocks;
}
static int attn16_slices(int64_t rows) {
    long long s = ceil_div(rows > 0 ? rows : 1, 4 * TS_RC);
    if (s > 6LL * sm_count()) s = 6LL * sm_count();     // several resident blocks per SM hide the chunk loads
    return (int)(s < 1 ? 1 : s);
}

struct Attn16Ws { size_t G, O, GR, wpart, p1, p2, total; int grid, slices; long long slice_rows; };
static Attn16Ws attn16_layout(int64_t B, int F, int A) {
    Attn16Ws L;
    const size_t rows = (size_t)B * F;
    L.grid = attn16_grid(B);
    L.slices = attn16_slices((int64_t)rows);
    L.slice_rows = ceil_div(ceil_div((int64_t)rows > 0 ? (int64_t)rows : 1, L.slices), TS_RC) * TS_RC;
    L.slices = (int)ceil_div((int64_t)rows > 0 ? (int64_t)rows : 1, L.slice_rows);
    size_t o = 0;
    auto take = [&](size_t floats) { size_t r = o; o += align_up(floats * 4, 256); return r; };
    L.G = take(rows * 3 * A); L.O = take(rows * A); L.GR = take(rows * 16);
    L.wpart = take((size_t)L.grid * A16_WARPS * 32);
    L.p1 = take((size_t)L.slices * ((size_t)3 * A * 16 + 3 * A));
    L.p2 = take((size_t)L.slices * ((size_t)16 * A + 16));
    L.total = o;
    return L;
}

size_t attn16_workspace_bytes(int64_t B, int F, int A) { return attn16_layout(B, F, A).total; }

int attn16_fwd(const float* x, int64_t B, int F, int A, int heads, int residual, const float* const* params, float* out,
               cudaStream_t st) {
    Attn16Args a;
    attn16_fill(a, x, nullptr, out, B, F, A, heads, residual, params);
    const size_t smem = attn16_smem_bytes(A, false);
    DFM_CHECK_CUDA(cudaFuncSetAttribute(attn16_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn16_fwd_kernel<<<attn16_grid(B), A16_WARPS * 32, smem, st>>>(a);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int attn16_bwd(const float* x, const float* g_out, int64_t B, int F, int A, int heads, int residual,
               const float* const* params, float* g_x, float* const* g_params, void* workspace, size_t workspace_bytes,
               cudaStream_t st) {
    const Attn16Ws L = attn16_layout(B, F, A);
    DFM_REQUIRE(workspace && workspace_bytes >= L.total, DFM_ERR_WORKSPACE, "dfm_attn_bwd: workspace %zu < %zu", workspace_bytes, L.total);
    char* ws = static_cast<char*>(workspace);
    Attn16Args a;
    attn16_fill(a, x, g_out, g_x, B, F, A, heads, residual, params);
    a.G = reinterpret_cast<float*>(ws + L.G); a.O = reinterpret_cast<float*>(ws + L.O);
    a.GR = reinterpret_cast<float*>(ws + L.GR); a.wpart = reinterpret_cast<float*>(ws + L.wpart);
    float* p1 = reinterpret_cast<float*>(ws + L.p1);
    float* p2 = reinterpret_cast<float*>(ws + L.p2);
    const size_t smem = attn16_smem_bytes(A, true);
    DFM_CHECK_CUDA(cudaFuncSetAttribute(attn16_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFM_CHECK_CUDA(cudaMemsetAsync(a.wpart, 0, (size_t)L.grid * A16_WARPS * 32 * 4, st));   // warps without a sample
    attn16_bwd_kernel<<<L.grid, A16_WARPS * 32, smem, st>>>(a);
    const long long rows = (long long)B * F;
    // [dWq | dWk | dWv] (3A x 16) = [gQ | gK | gV]^T x  (+ column sums = the bias gradients);  dWo (16 x A) = g_r^T O
    const size_t sm1 = (size_t)TS_RC * (3 * A + 16) * 4, sm2 = (size_t)TS_RC * (16 + A) * 4;
    if (3 * A <= 256) {
        DFM_CHECK_CUDA(cudaFuncSetAttribute(tsgemm_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
        tsgemm_kernel<8, 1><<<L.slices, 256, sm1, st>>>(a.G, 3 * A, 3 * A, x, 16, 16, rows, L.slice_rows, p1);
    } else {
        DFM_CHECK_CUDA(cudaFuncSetAttribute(tsgemm_kernel<12, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
        tsgemm_kernel<12, 1><<<L.slices, 256, sm1, st>>>(a.G, 3 * A, 3 * A, x, 16, 16, rows, L.slice_rows, p1);
    }
    tsgemm_kernel<1, 4><<<L.slices, 256, sm2, st>>>(a.GR, 16, 16, a.O, A, A, rows, L.slice_rows, p2);   // A <= 128
    SegOut s1, s2, s3;
    memset(&s1, 0, sizeof(s1)); memset(&s2, 0, sizeof(s2)); memset(&s3, 0, sizeof(s3));
    const int AD = A * 16;
    // order of g_params: Wq bq Wk bk Wv bv Wo bo gamma beta
    s1.n = 6;
    s1.dst[0] = g_params[0]; s1.off[0] = 0;          s1.len[0] = AD;
    s1.dst[1] = g_params[2]; s1.off[1] = AD;         s1.len[1] = AD;
    s1.dst[2] = g_params[4]; s1.off[2] = 2 * AD;     s1.len[2] = AD;
    s1.dst[3] = g_params[1]; s1.off[3] = 3 * AD;     s1.len[3] = A;
    s1.dst[4] = g_params[3]; s1.off[4] = 3 * AD + A; s1.len[4] = A;
    s1.dst[5] = g_params[5]; s1.off[5] = 3 * AD + 2 * A; s1.len[5] = A;
    const int len1 = 3 * AD + 3 * A;
    seg_reduce_kernel<<<(unsigned)ceil_div(len1, 128), 128, 0, st>>>(p1, L.slices, len1, s1);
    s2.n = 2;
    s2.dst[0] = g_params[6]; s2.off[0] = 0;  s2.len[0] = AD;
    s2.dst[1] = g_params[7]; s2.off[1] = AD; s2.len[1] = 16;
    seg_reduce_kernel<<<(unsigned)ceil_div(AD + 16, 128), 128, 0, st>>>(p2, L.slices, AD + 16, s2);
    if (residual) {
        s3.n = 2;
        s3.dst[0] = g_params[8]; s3.off[0] = 0;  s3.len[0] = 16;
        s3.dst[1] = g_params[9]; s3.off[1] = 16; s3.len[1] = 16;
        seg_reduce_kernel<<<1, 32, 0, st>>>(a.wpart, L.grid * A16_WARPS, 32, s3);
    }
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // namespace dfm
