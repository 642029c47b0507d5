// Field self-attention block (AttentionDeepFM), one fused shared-memory kernel per sample tile.
//
// Reference: deepfm/models/layers/attention.py:91-120 (_AttentionBlock.forward):
//     Q,K,V = Linear(D->A)(x);  per head softmax(Q K^T / sqrt(hd)) V;  W_out(A->D);
//     optional LayerNorm(out + x)  (eps 1e-5, affine).
// The whole block of one sample (F <= ~64 fields) lives in shared memory: x, Q, K, V, the
// (heads, F, F) probabilities, the merged heads and the output never touch HBM; only x is read
// and the (F, D) result written.  The backward recomputes the forward (nothing but x is saved),
// back-propagates inside shared memory, accumulates the parameter gradients of its samples in
// shared memory and writes ONE partial vector per block; a second kernel adds the partials in
// block order (deterministic, no float atomics).
#include "common.cuh"

namespace dfm {

struct AttnArgs {
    const float* x;
    const float* g_out;       // backward only
    float* out;               // forward: y ; backward: g_x
    long long B;
    int F, D, A, heads, hd, residual, spb, acc_global, nat;   // nat: natural-layout weight copies staged (backward)
    float eps, inv_scale;
    const float *Wq, *bq, *Wk, *bk, *Wv, *bv, *Wo, *bo, *gamma, *beta;
    float* partials;          // backward: (gridDim.x, NP)
};

struct AttnSmem {
    // weights, transposed + padded: WqT[d][a] at d*(A+4)+a ; WoT[a][d] at a*(D+4)+d  (rows stay 16-byte aligned)
    float *WqT, *WkT, *WvT, *WoT, *bq, *bk, *bv, *bo, *gamma, *beta;
    float *WqN, *WkN, *WvN, *WoN;    // backward: natural layouts Wq [A][D], Wo [D][A]
    // per-sample slots (spb of each)
    float *x, *q, *k, *v, *p, *o, *y;
    float *gy, *go, *gq, *gk, *gv;   // backward
    float* acc;                       // backward: parameter-gradient accumulators (NP)
};

__host__ __device__ inline int attn_np(int D, int A) { return 3 * (A * D + A) + D * A + D + 2 * D; }

__device__ __forceinline__ void attn_carve(float* base, const AttnArgs& a, bool bwd, AttnSmem& s) {
    const int F = a.F, D = a.D, A = a.A, spb = a.spb;
    float* p = base;
    auto take = [&](int n) { float* r = p; p += (n + 3) & ~3; return r; };
    s.WqT = take(D * (A + 4)); s.WkT = take(D * (A + 4)); s.WvT = take(D * (A + 4)); s.WoT = take(A * (D + 4));
    s.bq = take(A); s.bk = take(A); s.bv = take(A); s.bo = take(D); s.gamma = take(D); s.beta = take(D);
    s.x = take(spb * F * D); s.q = take(spb * F * (A + 4)); s.k = take(spb * F * (A + 4)); s.v = take(spb * F * (A + 4));
    s.p = take(spb * a.heads * F * F); s.o = take(spb * F * (A + 4)); s.y = take(spb * F * D);
    if (bwd) {
        s.gy = take(spb * F * D); s.go = take(spb * F * (A + 4)); s.gq = take(spb * F * (A + 4));
        s.gk = take(spb * F * (A + 4)); s.gv = take(spb * F * (A + 4));
        if (a.nat) { s.WqN = take(A * D); s.WkN = take(A * D); s.WvN = take(A * D); s.WoN = take(D * A); }
        s.acc = a.acc_global ? nullptr : take(attn_np(D, A));
    }
}

static size_t attn_smem_floats(int F, int D, int A, int heads, int spb, bool bwd, bool acc_global = false, bool nat = true) {
    auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
    size_t n = 3 * r4((size_t)D * (A + 4)) + r4((size_t)A * (D + 4)) + 3 * r4(A) + 3 * r4(D);
    n += 2 * r4((size_t)spb * F * D) + 4 * r4((size_t)spb * F * (A + 4)) + r4((size_t)spb * heads * F * F);
    if (bwd) n += r4((size_t)spb * F * D) + 4 * r4((size_t)spb * F * (A + 4)) + (nat ? 4 * r4((size_t)A * D) : 0) + (acc_global ? 0 : r4(attn_np(D, A)));
    return n;
}

__device__ __forceinline__ void attn_load_weights(const AttnArgs& a, const AttnSmem& s, bool bwd) {
    const int D = a.D, A = a.A, nt = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < A * D; i += nt) {
        const int r = i / D, c = i - r * D;     // Wq[r=a][c=d]
        s.WqT[c * (A + 4) + r] = __ldg(a.Wq + i);
        s.WkT[c * (A + 4) + r] = __ldg(a.Wk + i);
        s.WvT[c * (A + 4) + r] = __ldg(a.Wv + i);
    }
    for (int i = tid; i < D * A; i += nt) {
        const int r = i / A, c = i - r * A;     // Wo[r=d][c=a]
        s.WoT[c * (D + 4) + r] = __ldg(a.Wo + i);
    }
    if (bwd && a.nat) {
        for (int i = tid; i < A * D; i += nt) {
            s.WqN[i] = __ldg(a.Wq + i); s.WkN[i] = __ldg(a.Wk + i); s.WvN[i] = __ldg(a.Wv + i); s.WoN[i] = __ldg(a.Wo + i);
        }
    }
    for (int i = tid; i < A; i += nt) { s.bq[i] = __ldg(a.bq + i); s.bk[i] = __ldg(a.bk + i); s.bv[i] = __ldg(a.bv + i); }
    for (int i = tid; i < D; i += nt) {
        s.bo[i] = __ldg(a.bo + i);
        s.gamma[i] = a.residual ? __ldg(a.gamma + i) : 1.f;
        s.beta[i] = a.residual ? __ldg(a.beta + i) : 0.f;
    }
}


// Register-tiled shared-memory GEMM: every thread owns 4 x 4 tiles of C.
//   C[r*ldc + n] = init + sum_k A[r*sar + k*sak] * B[k*ldb + n],   r < R, n < N (N % 4 == 0, ldb % 4 == 0)
// INIT 0: zero, 1: bias[n], 2: C itself (accumulate).  One 128-bit B read and four broadcast A reads
// feed 16 FMAs (the plain per-element loops needed one shared-memory read per FMA).
template <int INIT>
__device__ __forceinline__ void sgemm_tile(const float* __restrict__ A, int sar, int sak, const float* __restrict__ B, int ldb,
                                           float* C, int ldc, int R, int N, int K, const float* __restrict__ bias) {
    const int tn_count = N >> 2, tr_count = (R + 3) >> 2;
    for (int tile = threadIdx.x; tile < tn_count * tr_count; tile += blockDim.x) {
        const int tn = tile % tn_count, tr = tile / tn_count;
        const int n0 = tn << 2, r0 = tr << 2;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (INIT == 1) acc[i][j] = bias[n0 + j];
                else if (INIT == 2) acc[i][j] = (r0 + i < R) ? C[(r0 + i) * ldc + n0 + j] : 0.f;
                else acc[i][j] = 0.f;
            }
        const float* a0 = A + (r0 + 0 < R ? r0 + 0 : R - 1) * sar;
        const float* a1 = A + (r0 + 1 < R ? r0 + 1 : R - 1) * sar;
        const float* a2 = A + (r0 + 2 < R ? r0 + 2 : R - 1) * sar;
        const float* a3 = A + (r0 + 3 < R ? r0 + 3 : R - 1) * sar;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 b4 = *reinterpret_cast<const float4*>(B + k * ldb + n0);
            const float av[4] = {a0[k * sak], a1[k * sak], a2[k * sak], a3[k * sak]};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (r0 + i < R) {
#pragma unroll
                for (int j = 0; j < 4; ++j) C[(r0 + i) * ldc + n0 + j] = acc[i][j];
            }
    }
}

// Phases 1-4 of the forward for `ns` samples already staged in s.x: fills q, k, v, p, o, y
// (y = W_out o + b_out, before the residual / LayerNorm).
__device__ __forceinline__ void attn_forward_core(const AttnArgs& a, const AttnSmem& s, int ns) {
    const int F = a.F, D = a.D, A = a.A, H = a.heads, hd = a.hd, nt = blockDim.x, tid = threadIdx.x;
    const int AS = A + 4;   // row stride of q/k/v/o (+4: rows of one head no longer share a bank, 16-byte aligned)
    // 1. Q, K, V
    const bool fast = (D % 4 == 0) && (A % 4 == 0);
    if (fast) {
        sgemm_tile<1>(s.x, D, 1, s.WqT, A + 4, s.q, AS, ns * F, A, D, s.bq);
        sgemm_tile<1>(s.x, D, 1, s.WkT, A + 4, s.k, AS, ns * F, A, D, s.bk);
        sgemm_tile<1>(s.x, D, 1, s.WvT, A + 4, s.v, AS, ns * F, A, D, s.bv);
    } else {
        for (int i = tid; i < ns * F * A; i += nt) {
            const int aa = i % A, row = i / A;          // row = s*F + f
            const float* xr = s.x + row * D;
            float q = s.bq[aa], k = s.bk[aa], v = s.bv[aa];
            for (int d = 0; d < D; ++d) {
                const float xv = xr[d];
                q = fmaf(xv, s.WqT[d * (A + 4) + aa], q);
                k = fmaf(xv, s.WkT[d * (A + 4) + aa], k);
                v = fmaf(xv, s.WvT[d * (A + 4) + aa], v);
            }
            s.q[row * AS + aa] = q; s.k[row * AS + aa] = k; s.v[row * AS + aa] = v;
        }
    }
    __syncthreads();
    // 2. scores + softmax, one thread per (sample, head, query row)
    for (int i = tid; i < ns * H * F; i += nt) {
        const int qi = i % F, h = (i / F) % H, sm = i / (F * H);
        const float* qr = s.q + (sm * F + qi) * AS + h * hd;
        float* pr = s.p + ((sm * H + h) * F + qi) * F;
        float mx = -INFINITY;
        for (int j = 0; j < F; ++j) {
            const float* kr = s.k + (sm * F + j) * AS + h * hd;
            float sc = 0.f;
            for (int c = 0; c < hd; ++c) sc = fmaf(qr[c], kr[c], sc);
            sc *= a.inv_scale;
            pr[j] = sc;
            mx = fmaxf(mx, sc);
        }
        float sum = 0.f;
        for (int j = 0; j < F; ++j) { const float e = expf(pr[j] - mx); pr[j] = e; sum += e; }
        const float inv = 1.f / sum;
        for (int j = 0; j < F; ++j) pr[j] *= inv;
    }
    __syncthreads();
    // 3. o = P V   (heads merged: column a belongs to head a / hd)
    for (int i = tid; i < ns * F * A; i += nt) {
        const int aa = i % A, row = i / A, qi = row % F, sm = row / F, h = aa / hd;
        const float* pr = s.p + ((sm * H + h) * F + qi) * F;
        float acc = 0.f;
        for (int j = 0; j < F; ++j) acc = fmaf(pr[j], s.v[(sm * F + j) * AS + aa], acc);
        s.o[row * AS + aa] = acc;
    }
    __syncthreads();
    // 4. y = o W_out^T + b_out
    if (fast) {
        sgemm_tile<1>(s.o, AS, 1, s.WoT, D + 4, s.y, D, ns * F, D, A, s.bo);
    } else {
        for (int i = tid; i < ns * F * D; i += nt) {
            const int d = i % D, row = i / D;
            const float* orow = s.o + row * AS;
            float acc = s.bo[d];
            for (int c = 0; c < A; ++c) acc = fmaf(orow[c], s.WoT[c * (D + 4) + d], acc);
            s.y[i] = acc;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
attn_fwd_kernel(const __grid_constant__ AttnArgs a) {
    extern __shared__ float smem[];
    AttnSmem s;
    attn_carve(smem, a, false, s);
    attn_load_weights(a, s, false);
    const int F = a.F, D = a.D, nt = blockDim.x, tid = threadIdx.x;
    const long long n_tiles = (a.B + a.spb - 1) / a.spb;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * a.spb;
        const int ns = (int)((a.B - b0 < a.spb) ? a.B - b0 : a.spb);
        __syncthreads();
        for (int i = tid; i < ns * F * D; i += nt) s.x[i] = __ldcs(a.x + b0 * F * D + i);
        __syncthreads();
        attn_forward_core(a, s, ns);
        if (a.residual) {   // LayerNorm(y + x), one thread per (sample, field) row, biased variance
            for (int r = tid; r < ns * F; r += nt) {
                float* yr = s.y + r * D;
                const float* xr = s.x + r * D;
                float mean = 0.f;
                for (int d = 0; d < D; ++d) { yr[d] += xr[d]; mean += yr[d]; }
                mean /= (float)D;
                float var = 0.f;
                for (int d = 0; d < D; ++d) { const float c = yr[d] - mean; var = fmaf(c, c, var); }
                const float rstd = rsqrtf(var / (float)D + a.eps);
                for (int d = 0; d < D; ++d) yr[d] = (yr[d] - mean) * rstd * s.gamma[d] + s.beta[d];
            }
            __syncthreads();
        }
        for (int i = tid; i < ns * F * D; i += nt) __stcs(a.out + b0 * F * D + i, s.y[i]);
    }
}

__global__ void __launch_bounds__(256)
attn_bwd_kernel(const __grid_constant__ AttnArgs a) {
    extern __shared__ float smem[];
    AttnSmem s;
    attn_carve(smem, a, true, s);
    attn_load_weights(a, s, true);
    const int F = a.F, D = a.D, A = a.A, H = a.heads, hd = a.hd, nt = blockDim.x, tid = threadIdx.x;
    const int NP = attn_np(D, A);
    const int AS = A + 4;
    // accumulator layout (shared memory, or this block's own row of the partials when that does not fit)
    if (a.acc_global) s.acc = a.partials + (size_t)blockIdx.x * NP;
    float* dWq = s.acc; float* dbq = dWq + A * D;
    float* dWk = dbq + A; float* dbk = dWk + A * D;
    float* dWv = dbk + A; float* dbv = dWv + A * D;
    float* dWo = dbv + A; float* dbo = dWo + D * A;
    float* dga = dbo + D; float* dbe = dga + D;
    for (int i = tid; i < NP; i += nt) s.acc[i] = 0.f;
    const long long n_tiles = (a.B + a.spb - 1) / a.spb;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * a.spb;
        const int ns = (int)((a.B - b0 < a.spb) ? a.B - b0 : a.spb);
        const int rows = ns * F;
        __syncthreads();
        for (int i = tid; i < rows * D; i += nt) {
            s.x[i] = __ldcs(a.x + b0 * F * D + i);
            s.gy[i] = __ldcs(a.g_out + b0 * F * D + i);
        }
        __syncthreads();
        attn_forward_core(a, s, ns);
        if (a.residual) {
            // (a) xhat -> s.y ; rstd kept in s.go[row] (free until step 4)
            for (int r = tid; r < rows; r += nt) {
                float* yr = s.y + r * D;
                const float* xr = s.x + r * D;
                float mean = 0.f;
                for (int d = 0; d < D; ++d) { yr[d] += xr[d]; mean += yr[d]; }
                mean /= (float)D;
                float var = 0.f;
                for (int d = 0; d < D; ++d) { const float c = yr[d] - mean; var = fmaf(c, c, var); }
                const float rstd = rsqrtf(var / (float)D + a.eps);
                for (int d = 0; d < D; ++d) yr[d] = (yr[d] - mean) * rstd;
                s.go[r] = rstd;
            }
            __syncthreads();
            // (b) d gamma / d beta, one thread per d, rows in order
            for (int d = tid; d < D; d += nt) {
                float ga = 0.f, be = 0.f;
                for (int r = 0; r < rows; ++r) { const float g = s.gy[r * D + d]; ga = fmaf(g, s.y[r * D + d], ga); be += g; }
                dga[d] += ga; dbe[d] += be;
            }
            __syncthreads();
            // (c) g_r = rstd * (gxh - mean(gxh) - xhat * mean(gxh * xhat)),  gxh = g * gamma
            for (int r = tid; r < rows; r += nt) {
                float* gr = s.gy + r * D;
                const float* xh = s.y + r * D;
                float m1 = 0.f, m2 = 0.f;
                for (int d = 0; d < D; ++d) { const float t = gr[d] * s.gamma[d]; m1 += t; m2 = fmaf(t, xh[d], m2); }
                m1 /= (float)D; m2 /= (float)D;
                const float rstd = s.go[r];
                for (int d = 0; d < D; ++d) gr[d] = rstd * (gr[d] * s.gamma[d] - m1 - xh[d] * m2);
            }
            __syncthreads();
        }
        // 4. W_out: dWo[d][a] += sum_rows gy[row][d] o[row][a] ; dbo ; go = gy W_out
        const bool fast = (D % 4 == 0) && (A % 4 == 0);
        if (fast) {
            sgemm_tile<2>(s.gy, 1, D, s.o, AS, dWo, A, D, A, rows, nullptr);
        } else {
            for (int i = tid; i < D * A; i += nt) {
                const int d = i / A, c = i - d * A;
                float acc = 0.f;
                for (int r = 0; r < rows; ++r) acc = fmaf(s.gy[r * D + d], s.o[r * AS + c], acc);
                dWo[i] += acc;
            }
        }
        for (int d = tid; d < D; d += nt) {
            float acc = 0.f;
            for (int r = 0; r < rows; ++r) acc += s.gy[r * D + d];
            dbo[d] += acc;
        }
        __syncthreads();   // rstd in s.go no longer needed
        if (fast && a.nat) {
            sgemm_tile<0>(s.gy, D, 1, s.WoN, A, s.go, AS, rows, A, D, nullptr);
        } else {
            for (int i = tid; i < rows * A; i += nt) {
                const int c = i % A, r = i / A;
                float acc = 0.f;
                for (int d = 0; d < D; ++d) acc = fmaf(s.gy[r * D + d], s.WoT[c * (D + 4) + d], acc);
                s.go[r * AS + c] = acc;
            }
        }
        __syncthreads();
        // 5. gv[j][a] = sum_i p[i][j] go[i][a]  (needs p: before p is overwritten)
        for (int i = tid; i < rows * A; i += nt) {
            const int c = i % A, row = i / A, j = row % F, sm = row / F, h = c / hd;
            float acc = 0.f;
            for (int qi = 0; qi < F; ++qi) acc = fmaf(s.p[((sm * H + h) * F + qi) * F + j], s.go[(sm * F + qi) * AS + c], acc);
            s.gv[row * AS + c] = acc;
        }
        __syncthreads();
        //    gs = p * (gp - sum_j gp p) / scale, in place over p ; gp[i][j] = sum_c go[i][c] v[j][c]
        for (int i = tid; i < ns * H * F; i += nt) {
            const int qi = i % F, h = (i / F) % H, sm = i / (F * H);
            float* pr = s.p + ((sm * H + h) * F + qi) * F;
            const float* gor = s.go + (sm * F + qi) * AS + h * hd;
            float dot = 0.f;
            for (int j = 0; j < F; ++j) {
                const float* vr = s.v + (sm * F + j) * AS + h * hd;
                float gp = 0.f;
                for (int c = 0; c < hd; ++c) gp = fmaf(gor[c], vr[c], gp);
                dot = fmaf(gp, pr[j], dot);
            }
            for (int j = 0; j < F; ++j) {
                const float* vr = s.v + (sm * F + j) * AS + h * hd;
                float gp = 0.f;
                for (int c = 0; c < hd; ++c) gp = fmaf(gor[c], vr[c], gp);
                pr[j] = pr[j] * (gp - dot) * a.inv_scale;
            }
        }
        __syncthreads();
        // 6. gq[i][a] = sum_j gs[i][j] k[j][a] ; gk[j][a] = sum_i gs[i][j] q[i][a]
        for (int i = tid; i < rows * A; i += nt) {
            const int c = i % A, row = i / A, f = row % F, sm = row / F, h = c / hd;
            const float* gs = s.p + (sm * H + h) * F * F;
            float aq = 0.f, ak = 0.f;
            for (int j = 0; j < F; ++j) {
                aq = fmaf(gs[f * F + j], s.k[(sm * F + j) * AS + c], aq);
                ak = fmaf(gs[j * F + f], s.q[(sm * F + j) * AS + c], ak);
            }
            s.gq[row * AS + c] = aq; s.gk[row * AS + c] = ak;
        }
        __syncthreads();
        // 7. dW{q,k,v}[a][d] += sum_rows g[row][a] x[row][d] ; biases ; g_x
        if (fast) {
            sgemm_tile<2>(s.gq, 1, AS, s.x, D, dWq, D, A, D, rows, nullptr);
            sgemm_tile<2>(s.gk, 1, AS, s.x, D, dWk, D, A, D, rows, nullptr);
            sgemm_tile<2>(s.gv, 1, AS, s.x, D, dWv, D, A, D, rows, nullptr);
        } else {
            for (int i = tid; i < A * D; i += nt) {
                const int c = i / D, d = i - c * D;
                float wq = 0.f, wk = 0.f, wv = 0.f;
                for (int r = 0; r < rows; ++r) {
                    const float xv = s.x[r * D + d];
                    wq = fmaf(s.gq[r * AS + c], xv, wq);
                    wk = fmaf(s.gk[r * AS + c], xv, wk);
                    wv = fmaf(s.gv[r * AS + c], xv, wv);
                }
                dWq[i] += wq; dWk[i] += wk; dWv[i] += wv;
            }
        }
        for (int c = tid; c < A; c += nt) {
            float q = 0.f, k = 0.f, v = 0.f;
            for (int r = 0; r < rows; ++r) { q += s.gq[r * AS + c]; k += s.gk[r * AS + c]; v += s.gv[r * AS + c]; }
            dbq[c] += q; dbk[c] += k; dbv[c] += v;
        }
        if (fast && a.nat) {
            // g_x = (residual ? g_r : 0) + gq Wq + gk Wk + gv Wv, accumulated in s.y (xhat is no longer needed)
            __syncthreads();
            for (int i = tid; i < rows * D; i += nt) s.y[i] = a.residual ? s.gy[i] : 0.f;
            __syncthreads();
            sgemm_tile<2>(s.gq, AS, 1, s.WqN, D, s.y, D, rows, D, A, nullptr);
            __syncthreads();
            sgemm_tile<2>(s.gk, AS, 1, s.WkN, D, s.y, D, rows, D, A, nullptr);
            __syncthreads();
            sgemm_tile<2>(s.gv, AS, 1, s.WvN, D, s.y, D, rows, D, A, nullptr);
            __syncthreads();
            for (int i = tid; i < rows * D; i += nt) __stcs(a.out + b0 * F * D + i, s.y[i]);
        } else {
            for (int i = tid; i < rows * D; i += nt) {
                const int d = i % D, r = i / D;
                float acc = a.residual ? s.gy[i] : 0.f;
                for (int c = 0; c < A; ++c) {
                    acc = fmaf(s.gq[r * AS + c], s.WqT[d * (A + 4) + c], acc);
                    acc = fmaf(s.gk[r * AS + c], s.WkT[d * (A + 4) + c], acc);
                    acc = fmaf(s.gv[r * AS + c], s.WvT[d * (A + 4) + c], acc);
                }
                __stcs(a.out + b0 * F * D + i, acc);
            }
        }
    }
    __syncthreads();
    if (!a.acc_global)
        for (int i = tid; i < NP; i += nt) a.partials[(size_t)blockIdx.x * NP + i] = s.acc[i];
}

struct AttnGradPtrs { float* p[10]; int n[10]; };

__global__ void attn_reduce_kernel(const float* __restrict__ partials, int n_blocks, int NP,
                                   const __grid_constant__ AttnGradPtrs gp, int residual) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NP) return;
    float acc = 0.f;
    for (int b = 0; b < n_blocks; ++b) acc += partials[(size_t)b * NP + i];
    int off = i;
    // accumulator order: Wq bq Wk bk Wv bv Wo bo gamma beta
    for (int t = 0; t < 10; ++t) {
        if (off < gp.n[t]) { if (gp.p[t]) gp.p[t][off] = acc; return; }
        off -= gp.n[t];
    }
}

static int attn_config(int64_t B, int F, int D, int A, int heads, bool bwd, int& spb, size_t& smem, int& grid,
                       int* acc_global = nullptr, int* nat = nullptr) {
    DFM_REQUIRE(F > 0 && D > 0 && A > 0 && heads > 0 && A % heads == 0, DFM_ERR_INVALID,
                "attn: need F, D, A > 0 and attention_dim %% num_heads == 0");
    const size_t budget = 200 * 1024;
    spb = 8;
    while (spb > 1 && attn_smem_floats(F, D, A, heads, spb, bwd) * 4 > budget) spb >>= 1;
    smem = attn_smem_floats(F, D, A, heads, spb, bwd) * 4;
    if (acc_global) *acc_global = 0;
    if (nat) *nat = 1;
    if (bwd && smem > budget && nat) {          // drop the natural-layout weight copies first
        *nat = 0;
        smem = attn_smem_floats(F, D, A, heads, spb, bwd, false, false) * 4;
    }
    if (bwd && smem > budget && acc_global) {   // then keep the parameter-gradient accumulators in global memory
        *acc_global = 1;
        smem = attn_smem_floats(F, D, A, heads, spb, bwd, true, nat ? *nat != 0 : true) * 4;
    }
    DFM_REQUIRE(smem <= 227 * 1024, DFM_ERR_UNSUPPORTED, "attn: F=%d D=%d A=%d needs %zu B shared memory per sample", F, D, A, smem);
    long long tiles = ceil_div(B > 0 ? B : 1, spb);
    grid = (int)(tiles < sm_count() ? tiles : sm_count());
    return DFM_OK;
}

static void attn_fill(AttnArgs& a, const float* x, const float* g_out, float* out, int64_t B, int F, int D, int A,
                      int heads, int residual, const float* const* p, int spb) {
    a.x = x; a.g_out = g_out; a.out = out; a.B = B; a.F = F; a.D = D; a.A = A; a.heads = heads; a.hd = A / heads;
    a.residual = residual; a.spb = spb; a.eps = 1e-5f; a.inv_scale = 1.f / sqrtf((float)(A / heads));
    a.Wq = p[0]; a.bq = p[1]; a.Wk = p[2]; a.bk = p[3]; a.Wv = p[4]; a.bv = p[5]; a.Wo = p[6]; a.bo = p[7];
    a.gamma = residual ? p[8] : nullptr; a.beta = residual ? p[9] : nullptr; a.partials = nullptr; a.acc_global = 0; a.nat = 0;
}

// attn16.cu: warp-per-sample kernels for D = 16, head dim 16, F <= 16 (BASELINE config 3)
bool attn16_supported(int F, int D, int A, int heads);
size_t attn16_workspace_bytes(int64_t B, int F, int A);
int attn16_fwd(const float* x, int64_t B, int F, int A, int heads, int residual, const float* const* params, float* out,
               cudaStream_t st);
int attn16_bwd(const float* x, const float* g_out, int64_t B, int F, int A, int heads, int residual,
               const float* const* params, float* g_x, float* const* g_params, void* workspace, size_t workspace_bytes,
               cudaStream_t st);

}  // namespace dfm

using namespace dfm;

extern "C" {

size_t dfm_attn_workspace_bytes(int64_t batch, int n_fields, int dim, int attention_dim, int heads) {
    int spb, grid; size_t smem;
    int ag = 0, nt_ = 1;
    if (attn_config(batch, n_fields, dim, attention_dim, heads, true, spb, smem, grid, &ag, &nt_) != DFM_OK) return 0;
    size_t bytes = (size_t)grid * attn_np(dim, attention_dim) * 4 + 256;
    if (attn16_supported(n_fields, dim, attention_dim, heads)) {
        const size_t b16 = attn16_workspace_bytes(batch, n_fields, attention_dim);
        if (b16 > bytes) bytes = b16;
    }
    return bytes;
}

int dfm_attn_fwd(const float* x, int64_t batch, int n_fields, int dim, int attention_dim, int heads,
                 int use_residual, const float* const* params, float* out, void* stream) {
    DFM_REQUIRE(params && batch >= 0, DFM_ERR_INVALID, "dfm_attn_fwd: null argument");
    for (int i = 0; i < (use_residual ? 10 : 8); ++i) DFM_REQUIRE(params[i], DFM_ERR_INVALID, "dfm_attn_fwd: parameter %d is null", i);
    int spb, grid; size_t smem;
    int rc = attn_config(batch, n_fields, dim, attention_dim, heads, false, spb, smem, grid);
    if (rc) return rc;
    if (batch == 0) return DFM_OK;
    DFM_REQUIRE(x && out, DFM_ERR_INVALID, "dfm_attn_fwd: null tensor");
    if (attn16_supported(n_fields, dim, attention_dim, heads) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0)
        return attn16_fwd(x, batch, n_fields, attention_dim, heads, use_residual, params, out, static_cast<cudaStream_t>(stream));
    AttnArgs a;
    attn_fill(a, x, nullptr, out, batch, n_fields, dim, attention_dim, heads, use_residual, params, spb);
    if (smem > 48 * 1024) DFM_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_attn_bwd(const float* x, const float* g_out, int64_t batch, int n_fields, int dim, int attention_dim,
                 int heads, int use_residual, const float* const* params, float* g_x, float* const* g_params,
                 void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(params && g_params && batch >= 0, DFM_ERR_INVALID, "dfm_attn_bwd: null argument");
    for (int i = 0; i < (use_residual ? 10 : 8); ++i)
        DFM_REQUIRE(params[i] && g_params[i], DFM_ERR_INVALID, "dfm_attn_bwd: parameter %d is null", i);
    int spb, grid; size_t smem;
    int acc_global = 0, nat = 1;
    int rc = attn_config(batch, n_fields, dim, attention_dim, heads, true, spb, smem, grid, &acc_global, &nat);
    if (rc) return rc;
    const int D = dim, A = attention_dim, NP = attn_np(D, A);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (batch > 0 && attn16_supported(n_fields, dim, attention_dim, heads) && x && g_out && g_x &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g_out) | reinterpret_cast<uintptr_t>(g_x)) & 15u) == 0)
        return attn16_bwd(x, g_out, batch, n_fields, attention_dim, heads, use_residual, params, g_x, g_params, workspace,
                          workspace_bytes, st);
    AttnGradPtrs gp;
    const int sizes[10] = {A * D, A, A * D, A, A * D, A, D * A, D, D, D};
    for (int i = 0; i < 10; ++i) { gp.n[i] = sizes[i]; gp.p[i] = (i < 8 || use_residual) ? g_params[i] : nullptr; }
    if (batch == 0) grid = 0;
    DFM_REQUIRE(workspace && workspace_bytes >= (size_t)(grid > 0 ? grid : 1) * NP * 4, DFM_ERR_WORKSPACE, "dfm_attn_bwd: workspace too small");
    if (batch > 0) {
        DFM_REQUIRE(x && g_out && g_x, DFM_ERR_INVALID, "dfm_attn_bwd: null tensor");
        AttnArgs a;
        attn_fill(a, x, g_out, g_x, batch, n_fields, dim, attention_dim, heads, use_residual, params, spb);
        a.partials = static_cast<float*>(workspace);
        a.acc_global = acc_global;
        a.nat = nat;
        if (smem > 48 * 1024) DFM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attn_bwd_kernel<<<grid, 256, smem, st>>>(a);
    }
    attn_reduce_kernel<<<(unsigned)ceil_div(NP, 256), 256, 0, st>>>(static_cast<const float*>(workspace), grid, NP, gp, use_residual);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
