// CIN layer forward on the 5th-generation tensor cores (tcgen05, TF32 inputs, FP32 accumulate).
//
// The layer is the GEMM  C[(b,d)][l] = sum_k Z[(b,d)][k] * W[l][k],  k = h*F + f,
// Z[(b,d)][k] = hidden[b,h,d] * x0[b,f,d]   (reference: deepfm/models/layers/cin.py:84-91).
// Z never exists in memory: 8 producer warps synthesise 128 x 32 tiles of it straight into the
// 128B-swizzled K-major shared-memory layout a UMMA descriptor expects (one thread per GEMM row,
// the row's x0 values parked in shared memory, the hidden value streamed), while the same warps
// stage the matching W tile; one elected thread of a ninth warp issues
// `tcgen05.mma.cta_group::1.kind::tf32` (M = 128, N = L, K = 8 per instruction) into two TMEM
// accumulators (two 128-row tiles share every W tile); `tcgen05.commit` hands shared-memory stages
// back to the producers through mbarriers.  The epilogue reads the accumulators with
// `tcgen05.ld.32x32b`, adds the bias, applies ReLU and writes the (B, L, D) activation.
// Persistent grid: one CTA per SM walks 256-row super-tiles.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dfm {
namespace tc {

constexpr int ROWS = 256;        // GEMM rows per CTA super-tile (2 UMMA tiles of 128)
constexpr int KB = 32;           // k per stage: 32 tf32 = one 128-byte swizzle row
constexpr int NSTAGE = 4;
constexpr int PRODUCER_WARPS = 8;
constexpr int THREADS = (PRODUCER_WARPS + 1) * 32;

struct CinTcArgs {
    const float* x0; long long x_bs;       // x0[b*x_bs + f*D + d]
    const float* hid; long long h_bs;      // hidden[b*h_bs + h*D + d]
    const float* wpad;                     // (Np, Kp) zero-padded weight, row-major
    const float* bias;                     // (L)
    float* act;                            // (B, L, D)
    long long M;                           // B * D rows
    int F, FP, H, D, L, Np, Kp;            // FP: F padded to a multiple of 4 (k = h*FP + f)
    uint32_t tmem_cols;
    int nstage;                            // ring depth (<= NSTAGE)
};

// (L, H*F) -> (Np, Kp) zero padded, column k' = h*FP + f
__global__ void cin_pad_w_kernel(const float* __restrict__ w, int L, int H, int F, int FP, int Np, int Kp,
                                 float* __restrict__ out) {
    const long long n = (long long)Np * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i / Kp), k = (int)(i - (long long)l * Kp);
        const int h = k / FP, f = k - h * FP;
        out[i] = (l < L && h < H && f < F) ? __ldg(w + (size_t)l * H * F + h * F + f) : 0.f;
    }
}

// A_TMEM = true: the synthesised A tiles go registers -> tensor memory (tcgen05.st) and the MMA reads A from
// TMEM, B from shared memory: shared-memory traffic per k-block drops from 144 KB to 80 KB (the SS form
// at N = 128 is shared-memory-bandwidth bound).  Needs 2*Np + NSTAGE*64 <= 512 TMEM columns (Np <= 128).
// SFP (static FP): 40 (F = 37..40, the Criteo shapes) or 16 (F <= 16, the ML-100K shapes) -- with A_TMEM the thread's whole
// x0 row then lives in REGISTERS and the (h, f) of every k of a k-block is a compile-time constant of the k-block's
// position in the period lcm(32, FP) / 32 (5 k-blocks for FP = 40, 1 for FP = 16): a k-block costs the producer 32 FMUL +
// one tcgen05.st instead of 8 LDS.128 + ~150 instructions of index arithmetic and selects (267 -> ~60 warp instructions per
// warp per k-block against a 512-cycle MMA budget).  SFP = 0: the generic loop (x0 row in shared memory).
template <bool A_TMEM, int SFP>
__global__ void __launch_bounds__(THREADS, 1)
cin_tc_fwd_kernel(const __grid_constant__ CinTcArgs a, const __grid_constant__ CUtensorMap wmap) {
    extern __shared__ unsigned char smem_raw[];
    // the 128-byte swizzle is a function of the absolute shared address: tiles must be 1024-B aligned
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // carve: [stage A tiles: 2 x 16 KB][stage B tile: Np x 128 B] x NSTAGE, then x0 rows, barriers
    const int a_bytes = A_TMEM ? 0 : 2 * 128 * 128, b_bytes = a.Np * 128;
    const int stage_bytes = (a_bytes + b_bytes + 1023) & ~1023;
    unsigned char* stage0 = smem;
    const int FPS = a.FP + 4;              // row stride (floats): 128-bit reads of 8 lanes hit 8 distinct bank groups
    float* s_x0 = reinterpret_cast<float*>(smem + (size_t)a.nstage * stage_bytes);          // [ROWS][FPS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_x0 + (size_t)ROWS * FPS);
    uint64_t* full = bars;                 // [NSTAGE] producers -> MMA
    uint64_t* empty = bars + NSTAGE;       // [NSTAGE] MMA (commit) -> producers
    uint64_t* acc_full = bars + 2 * NSTAGE;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full + s, PRODUCER_WARPS + 1); mbar_init(empty + s, 1); }
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wmap)) : "memory");
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, PRODUCER_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PRODUCER_WARPS) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_kb = a.Kp / KB;
    const long long n_super = (a.M + ROWS - 1) / ROWS;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.Np >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_col0 = (uint32_t)(2 * a.Np);      // TMEM columns of the A-tile ring (A_TMEM)

    if (warp < PRODUCER_WARPS) {
        // ------------------------------------------------------------------ producers + epilogue
        const int r = threadIdx.x;                   // row inside the super-tile
        const int tile = r >> 7, row = r & 127;
        uint32_t s = 0, ph = 0;                      // ring position and phase of the next k-block
        uint32_t acc_phase = 0;
        for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
            const long long m = st * ROWS + r;
            const bool live = m < a.M;
            const long long b = live ? m / a.D : 0;
            const int d = live ? (int)(m - b * a.D) : 0;
            const float* xrow = a.x0 + b * a.x_bs + d;
            const float* hrow = a.hid + b * a.h_bs + d;
            float* xr = s_x0 + (size_t)r * FPS;          // this thread's x0 row, zero padded to FP
            if (!(A_TMEM && SFP > 0))
            for (int f0 = 0; f0 < a.FP; f0 += 8) {       // 8 independent loads in flight, then 2 STS.128
                float t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = (live && f0 + u < a.F) ? __ldg(xrow + (size_t)(f0 + u) * a.D) : 0.f;
                *reinterpret_cast<float4*>(xr + f0) = make_float4(t[0], t[1], t[2], t[3]);
                if (f0 + 4 < a.FP) *reinterpret_cast<float4*>(xr + f0 + 4) = make_float4(t[4], t[5], t[6], t[7]);
            }
            if (A_TMEM && SFP > 0) {
                constexpr int FPc = SFP > 0 ? SFP : 32;
                constexpr int PER = FPc == 40 ? 5 : 1;                 // k-blocks per period
                constexpr int HPP = PER * 32 / FPc;                    // h values per period (4 for FP = 40, 2 for FP = 16)
                float xreg[FPc];
#pragma unroll
                for (int f = 0; f < FPc; ++f) xreg[f] = (live && f < a.F) ? __ldg(xrow + (size_t)f * a.D) : 0.f;
                auto ldh = [&](int hh) { return (live && hh < a.H) ? __ldg(hrow + (size_t)hh * a.D) : 0.f; };
                float hw[HPP], hnext[HPP];
#pragma unroll
                for (int i = 0; i < HPP; ++i) hw[i] = ldh(i);
                for (int kb0 = 0, hb = 0; kb0 < n_kb; kb0 += PER, hb += HPP) {
#pragma unroll
                    for (int i = 0; i < HPP; ++i) hnext[i] = ldh(hb + HPP + i);      // next period's h values, one period ahead
#pragma unroll
                    for (int pp = 0; pp < PER; ++pp) {
                        if (kb0 + pp >= n_kb) break;                   // uniform over the CTA
                        mbar_wait(empty + s, ph ^ 1u);
                        if (r == 0) {
                            mbar_arrive_expect_tx(full + s, (uint32_t)a.Np * 128u);
                            tma_load_2d(stage0 + (size_t)s * stage_bytes + a_bytes, &wmap, (kb0 + pp) * KB, 0, full + s);
                        }
                        float z[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            constexpr int dummy = 0; (void)dummy;
                            const int kk = pp * 32 + k;                // position inside the period: compile-time after unrolling
                            z[k] = hw[kk / FPc] * xreg[kk % FPc];
                        }
                        tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + a_col0 + s * 64 + tile * 32, z);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full + s);
                        if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
                    }
#pragma unroll
                    for (int i = 0; i < HPP; ++i) hw[i] = hnext[i];
                }
            } else {
            // hidden values of the current and the next two h (a k-block of 32 spans at most 3: FP >= 16)
            int h = 0, f4 = 0;
            auto ldh = [&](int hh) { return (live && hh < a.H) ? __ldg(hrow + (size_t)hh * a.D) : 0.f; };
            float hw0 = ldh(0), hw1 = ldh(1), hw2 = ldh(2), hw3 = ldh(3), hw4 = ldh(4);   // hw3/hw4: prefetched, not yet needed
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(empty + s, ph ^ 1u);                          // stage free (first pass: immediately)
                if (r == 0) {   // W tile: one TMA box (32 x Np floats), lands swizzled, completes on `full`
                    mbar_arrive_expect_tx(full + s, (uint32_t)a.Np * 128u);
                    tma_load_2d(stage0 + (size_t)s * stage_bytes + a_bytes, &wmap, kb * KB, 0, full + s);
                }
                unsigned char* sa = stage0 + (size_t)s * stage_bytes + tile * (128 * 128) + row * 128;
                // 8 chunks of 4 consecutive f of one h: all 8 LDS.128 first, then 4 FMUL + STS.128 each
                float4 x4[8];
                int wc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    int fc = f4 + 4 * c, w = 0;
                    if (fc >= a.FP) { fc -= a.FP; w = 1; }
                    if (fc >= a.FP) { fc -= a.FP; w = 2; }
                    x4[c] = *reinterpret_cast<const float4*>(xr + fc);
                    wc[c] = w;
                }
                if (A_TMEM) {
                    float z[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float hv = wc[c] == 0 ? hw0 : (wc[c] == 1 ? hw1 : hw2);
                        z[4 * c] = hv * x4[c].x; z[4 * c + 1] = hv * x4[c].y; z[4 * c + 2] = hv * x4[c].z; z[4 * c + 3] = hv * x4[c].w;
                    }
                    // lane = row inside the 128-row tile, columns = the 32 k of this stage
                    tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + a_col0 + s * 64 + tile * 32, z);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float hv = wc[c] == 0 ? hw0 : (wc[c] == 1 ? hw1 : hw2);
                        *reinterpret_cast<float4*>(sa + ((c ^ (row & 7)) << 4)) =
                            make_float4(hv * x4[c].x, hv * x4[c].y, hv * x4[c].z, hv * x4[c].w);
                    }
                }
                int fe = f4 + KB, nw = 0;
                if (fe >= a.FP) { fe -= a.FP; nw = 1; }
                if (fe >= a.FP) { fe -= a.FP; nw = 2; }
                f4 = fe; h += nw;
                // slide the window; the freshly issued loads land in slots that are not read for >= 1 k-block
                if (nw == 1) { hw0 = hw1; hw1 = hw2; hw2 = hw3; hw3 = hw4; hw4 = ldh(h + 4); }
                else if (nw == 2) { hw0 = hw2; hw1 = hw3; hw2 = hw4; hw3 = ldh(h + 3); hw4 = ldh(h + 4); }
                if (A_TMEM) tc_fence_before();                           // tcgen05.st (waited) ordered before the arrive
                else fence_proxy_async();                                // generic-proxy writes -> async proxy (UMMA)
                __syncwarp();
                if (lane == 0) mbar_arrive(full + s);
                if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
            }
            }   // generic producer loop
            // ---- epilogue: TMEM -> registers -> bias + ReLU -> act[b][l][d]
            mbar_wait(acc_full, acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tile * a.Np);
            float* out = a.act + (size_t)b * a.L * a.D + d;
            for (int c0 = 0; c0 < a.Np; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                if (live) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int l = c0 + i;
                        if (l < a.L) out[(size_t)l * a.D] = fmaxf(v[i] + __ldg(a.bias + l), 0.f);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        uint32_t s = 0, ph = 0, acc_phase = 0;
        for (long long st = blockIdx.x; st < n_super; st += gridDim.x) {
            mbar_wait(acc_empty, acc_phase ^ 1u);                       // epilogue drained the accumulators
            acc_phase ^= 1u;
            tc_fence_after();
            for (int kb = 0; kb < n_kb; ++kb) {
                mbar_wait(full + s, ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(stage0 + (size_t)s * stage_bytes);
                const uint64_t db = make_desc(sa + a_bytes);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const uint64_t da = make_desc(sa + t * (128 * 128));
#pragma unroll
                    for (int k = 0; k < KB / 8; ++k) {                   // K = 8 tf32 = 32 B (8 TMEM columns) per instruction
                        if (A_TMEM)
                            umma_tf32_ts(tmem_base + (uint32_t)(t * a.Np), tmem_base + a_col0 + s * 64 + t * 32 + k * 8,
                                         db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                        else
                            umma_tf32(tmem_base + (uint32_t)(t * a.Np), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                                      (kb | k) ? 1u : 0u);
                    }
                }
                umma_commit(empty + s);                                  // stage reusable once these MMAs retire
                if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
            }
            umma_commit(acc_full);                                       // accumulators complete
        }
    }
    __syncthreads();
    if (warp == PRODUCER_WARPS) tmem_dealloc(tmem_base, a.tmem_cols);
}

}  // namespace tc

// One CIN layer forward on tcgen05.  wpad: workspace of Np*Kp floats.
int cin_layer_fwd_tc(const float* x0, long long x_bs, const float* hid, long long h_bs, const float* w,
                     const float* bias, float* act, long long B, int F, int H, int D, int L, float* wpad,
                     cudaStream_t st) {
    using namespace tc;
    const int FP = F <= 16 ? 16 : (F + 3) & ~3;     // k = h*FP + f; a 32-wide k-block then spans <= 3 values of h
    const int Np = (L + 15) & ~15, Kp = (H * FP + KB - 1) / KB * KB;
    DFM_REQUIRE(Np <= 256, DFM_ERR_UNSUPPORTED, "cin tcgen05: layer size %d > 256", L);
    long long pb = ceil_div((long long)Np * Kp, 256);
    if (pb > 4LL * sm_count()) pb = 4LL * sm_count();
    cin_pad_w_kernel<<<(unsigned)pb, 256, 0, st>>>(w, L, H, F, FP, Np, Kp, wpad);
    CinTcArgs a;
    a.x0 = x0; a.x_bs = x_bs; a.hid = hid; a.h_bs = h_bs; a.wpad = wpad; a.bias = bias; a.act = act;
    a.M = B * D; a.F = F; a.FP = FP; a.H = H; a.D = D; a.L = L; a.Np = Np; a.Kp = Kp;
    const bool a_tmem = 2 * Np + NSTAGE * 64 <= 512;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * Np + (a_tmem ? NSTAGE * 64 : 0))) cols <<= 1;
    a.tmem_cols = cols;
    const int stage_bytes = ((a_tmem ? 0 : 2 * 128 * 128) + Np * 128 + 1023) & ~1023;
    const size_t fixed = (size_t)ROWS * (FP + 4) * 4 + (2 * NSTAGE + 2) * 8 + 16 + 1024;
    int nstage = NSTAGE;
    while (nstage > 2 && (size_t)nstage * stage_bytes + fixed > 227 * 1024) --nstage;
    a.nstage = nstage;
    const size_t smem = (size_t)nstage * stage_bytes + fixed;
    DFM_REQUIRE(smem <= 227 * 1024, DFM_ERR_UNSUPPORTED, "cin tcgen05: %d fields need %zu B shared memory", F, smem);
    DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_tc_fwd_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_tc_fwd_kernel<true, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_tc_fwd_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_tc_fwd_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = ceil_div(a.M, ROWS);
    if (grid > sm_count()) grid = sm_count();
    // tensor map of the padded weight: (Kp inner, Np outer) fp32, box 32 x Np, 128-byte swizzle
    CUtensorMap wmap;
    int rc = make_tmap_2d(&wmap, wpad, Np, Kp, Np);
    if (rc) return rc;
    const bool stat = getenv("DFM_CIN_GENERIC") == nullptr;
    if (a_tmem && stat && FP == 40) cin_tc_fwd_kernel<true, 40><<<(unsigned)grid, THREADS, smem, st>>>(a, wmap);
    else if (a_tmem && stat && FP == 16) cin_tc_fwd_kernel<true, 16><<<(unsigned)grid, THREADS, smem, st>>>(a, wmap);
    else if (a_tmem) cin_tc_fwd_kernel<true, 0><<<(unsigned)grid, THREADS, smem, st>>>(a, wmap);
    else cin_tc_fwd_kernel<false, 0><<<(unsigned)grid, THREADS, smem, st>>>(a, wmap);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

size_t cin_tc_wpad_floats(int F, int Hmax, int Lmax) {
    const int Np = (Lmax + 15) & ~15;
    const long long Kp = ((long long)Hmax * (F <= 16 ? 16 : (F + 3) & ~3) + tc::KB - 1) / tc::KB * tc::KB;
    return (size_t)Np * Kp;
}

}  // namespace dfm
