// CIN backward w.r.t. the layer inputs on tcgen05 (TF32 inputs, FP32 accumulate).
//
// With g_pre = relu'(act) * upstream (rows (b,d), columns l), the gradient of the outer-product operand is
//     gz[(b,d)][k] = sum_l g_pre[(b,d)][l] * W[l][k],        k = h*F + f            (a plain GEMM, N = K)
// and the two input gradients are per-row contractions of gz:
//     g_hidden[(b,d)][h] = sum_f gz[(b,d)][h,f] * x0[(b,d)][f],   g_x0[(b,d)][f] += sum_h gz[(b,d)][h,f] * hidden[(b,d)][h]
// gz (as large as the never-materialised outer product) stays in tensor memory: the kernel computes
// it 128 rows x (HT*FP) columns at a time -- HT whole values of h per tile, so column c of a tile is
// f = c % FP, h = tile*HT + c / FP at COMPILE time -- and four epilogue warps (thread = row = TMEM lane)
// contract each tile against the row's x0 / hidden values held in registers while the next tile's
// MMAs run into the second accumulator.  Both operands are real tensors, staged by TMA:
//     A = g_pre^T  (B*D, Lp)  row-major (written in that layout by the g_pre kernel),
//     B = W^T pad  (Hp*FP, Lp) row-major, row k' = h*FP + f.
#include "tc_common.cuh"

namespace dfm {
namespace tc {

constexpr int BW_NSTAGE = 3;
constexpr int BW_THREADS = 6 * 32;     // 4 epilogue warps, 1 TMA warp, 1 MMA warp

__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// FP consecutive TMEM columns of this thread's lane -> registers (FP multiple of 8, <= 64)
template <int FP>
__device__ __forceinline__ void tmem_ld_row(uint32_t addr, float* v) {
    int c = 0;
#pragma unroll
    for (; c + 16 <= FP; c += 16) tmem_ld16(addr + c, v + c);
    if (FP % 16) tmem_ld8(addr + c, v + c);
    tmem_ld_wait();
}

struct BwdDataArgs {
    const float* x0; long long x_bs;       // x0[b*x_bs + f*D + d]
    const float* hid; long long h_bs;      // hidden[b*h_bs + h*D + d]
    float* g_hid; long long gh_bs;         // out[b*gh_bs + h*D + d]
    float* g_x0;                           // (B, F, D), accumulated into
    long long M;                           // B * D
    int F, H, D, Lp, n_tiles;              // n_tiles = ceil(H / HT)
    int gh_accumulate;                     // layer 0: hidden is x0 itself, g_hid is added to g_x0
    int nstage;
};

template <int FP>
__global__ void __launch_bounds__(BW_THREADS, 1)
cin_tc_bwd_data_kernel(const __grid_constant__ BwdDataArgs a, const __grid_constant__ CUtensorMap amap,
                       const __grid_constant__ CUtensorMap bmap) {
    constexpr int HT = 256 / FP;           // values of h per N tile
    constexpr int NT = HT * FP;            // MMA N (240 or 256)
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int n_kb = a.Lp / 32;
    const int a_bytes = n_kb * 128 * 128;                       // g_pre^T rows of this M tile, all of L
    const int b_stage = (NT * 128 + 1023) & ~1023;
    unsigned char* sA = smem;
    unsigned char* sB = smem + a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)a.nstage * b_stage);
    uint64_t* a_full = bars;               // TMA -> MMA (A tile resident)
    uint64_t* a_empty = bars + 1;          // MMA -> TMA (all MMAs of the M tile retired)
    uint64_t* b_full = bars + 2;           // [BW_NSTAGE]
    uint64_t* b_empty = b_full + BW_NSTAGE;
    uint64_t* acc_full = b_empty + BW_NSTAGE;   // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < BW_NSTAGE; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&amap)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&bmap)) : "memory");
    }
    if (warp == 5) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long n_mtiles = (a.M + 127) / 128;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);

    if (warp < 4) {
        // ------------------------------------------------------------ epilogue: thread = row = TMEM lane
        const int r = threadIdx.x;
        uint32_t accph[2] = {0u, 0u};
        uint32_t jt = 0;                                       // running N-tile counter (selects the accumulator)
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
            const long long m = mt * 128 + r;
            const bool live = m < a.M;
            const long long b = live ? m / a.D : 0;
            const int d = live ? (int)(m - b * a.D) : 0;
            const float* xrow = a.x0 + b * a.x_bs + d;
            const float* hrow = a.hid + b * a.h_bs + d;
            float xr[FP], gx[FP];
#pragma unroll
            for (int f = 0; f < FP; ++f) { xr[f] = (live && f < a.F) ? __ldg(xrow + (size_t)f * a.D) : 0.f; gx[f] = 0.f; }
            for (int j = 0; j < a.n_tiles; ++j, ++jt) {
                const uint32_t ab = jt & 1u;
                float hv[HT];
#pragma unroll
                for (int t = 0; t < HT; ++t) {
                    const int h = j * HT + t;
                    hv[t] = (live && h < a.H) ? __ldg(hrow + (size_t)h * a.D) : 0.f;
                }
                mbar_wait(acc_full + ab, accph[ab]);
                accph[ab] ^= 1u;
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + ab * 256u;
#pragma unroll
                for (int t = 0; t < HT; ++t) {
                    float gz[FP];
                    tmem_ld_row<FP>(taddr + t * FP, gz);
                    float gh = 0.f;
#pragma unroll
                    for (int f = 0; f < FP; ++f) {
                        gh = fmaf(gz[f], xr[f], gh);
                        gx[f] = fmaf(gz[f], hv[t], gx[f]);
                    }
                    const int h = j * HT + t;
                    if (live && h < a.H) {
                        float* o = a.g_hid + b * a.gh_bs + (size_t)h * a.D + d;
                        *o = a.gh_accumulate ? *o + gh : gh;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + ab);
            }
            if (live) {
                float* o = a.g_x0 + b * (long long)a.F * a.D + d;
#pragma unroll
                for (int f = 0; f < FP; ++f)
                    if (f < a.F) o[(size_t)f * a.D] += gx[f];
            }
        }
    } else if (warp == 4 && lane == 0) {
        // ------------------------------------------------------------ TMA producer
        uint32_t s = 0, ph = 0, aph = 0;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
            mbar_wait(a_empty, aph ^ 1u);
            aph ^= 1u;
            mbar_arrive_expect_tx(a_full, (uint32_t)a_bytes);
            for (int kb = 0; kb < n_kb; ++kb)
                tma_load_2d(sA + (size_t)kb * 128 * 128, &amap, kb * 32, (int)(mt * 128), a_full);
            for (int j = 0; j < a.n_tiles; ++j)
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(b_empty + s, ph ^ 1u);
                    mbar_arrive_expect_tx(b_full + s, (uint32_t)NT * 128u);
                    tma_load_2d(sB + (size_t)s * b_stage, &bmap, kb * 32, j * NT, b_full + s);
                    if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
                }
        }
    } else if (warp == 5 && lane == 0) {
        // ------------------------------------------------------------ MMA issuer
        uint32_t s = 0, ph = 0, aph = 0, accph[2] = {0u, 0u}, jt = 0;
        for (long long mt = blockIdx.x; mt < n_mtiles; mt += gridDim.x) {
            mbar_wait(a_full, aph);
            aph ^= 1u;
            for (int j = 0; j < a.n_tiles; ++j, ++jt) {
                const uint32_t ab = jt & 1u;
                mbar_wait(acc_empty + ab, accph[ab] ^ 1u);
                accph[ab] ^= 1u;
                tc_fence_after();
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(b_full + s, ph);
                    tc_fence_after();
                    const uint64_t da = make_desc(smem_u32(sA + (size_t)kb * 128 * 128));
                    const uint64_t db = make_desc(smem_u32(sB + (size_t)s * b_stage));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_tf32(tmem_base + ab * 256u, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(b_empty + s);
                    if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
                }
                umma_commit(acc_full + ab);
            }
            umma_commit(a_empty);
        }
    }
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// g_pre (B, L, D) -> g_pre^T (B*D, Lp), zero padded: one block per sample, through shared memory
__global__ void __launch_bounds__(256)
cin_gpre_transpose_kernel(const float* __restrict__ gp, long long B, int L, int D, int Lp, float* __restrict__ gT) {
    extern __shared__ float tile[];                      // [L][D + 1]
    const long long b = blockIdx.x;
    const float* src = gp + b * L * D;
    for (int i = threadIdx.x; i < L * D; i += 256) { const int l = i / D, d = i - l * D; tile[l * (D + 1) + d] = __ldg(src + i); }
    __syncthreads();
    float* dst = gT + b * D * Lp;
    for (int i = threadIdx.x; i < D * Lp; i += 256) {
        const int d = i / Lp, l = i - d * Lp;
        dst[i] = l < L ? tile[l * (D + 1) + d] : 0.f;
    }
}

// W (L, H*F) -> W^T pad (Hp*FP, Lp): row k' = h*FP + f, column l
__global__ void cin_wt_pad_kernel(const float* __restrict__ w, int L, int H, int F, int FP, int Hp, int Lp,
                                  float* __restrict__ out) {
    const long long n = (long long)Hp * FP * Lp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % Lp);
        const int k = (int)(i / Lp);
        const int h = k / FP, f = k - h * FP;
        out[i] = (l < L && h < H && f < F) ? __ldg(w + (size_t)l * H * F + h * F + f) : 0.f;
    }
}

template <int FP>
static int launch_bwd_data(const BwdDataArgs& a, const CUtensorMap& amap, const CUtensorMap& bmap, size_t smem, int grid,
                           cudaStream_t st) {
    DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_tc_bwd_data_kernel<FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cin_tc_bwd_data_kernel<FP><<<grid, BW_THREADS, smem, st>>>(a, amap, bmap);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // namespace tc

int cin_tc_fp(int F) { return F <= 16 ? 16 : F <= 24 ? 24 : F <= 32 ? 32 : F <= 40 ? 40 : F <= 48 ? 48 : F <= 64 ? 64 : 0; }

// scratch floats: g_pre^T (B*D*Lp) + W^T pad (Hp*FP*Lp)
size_t cin_tc_bwd_scratch_floats(long long B, int F, int D, int Hmax, int Lmax) {
    const int FP = cin_tc_fp(F);
    if (!FP) return 0;
    const int HT = 256 / FP, Lp = (Lmax + 31) & ~31, Hp = (Hmax + HT - 1) / HT * HT;
    return (size_t)B * D * Lp + (size_t)Hp * FP * Lp + 64;
}

// d/d hidden and d/d x0 of one CIN layer on tcgen05.  Returns DFM_ERR_UNSUPPORTED if the shape is outside
// the instantiated tiles (the caller then uses the fp32 CUDA-core path).
int cin_layer_bwd_data_tc(const float* g_pre, const float* x0, long long x_bs, const float* hid, long long h_bs,
                          const float* w, float* g_hid, long long gh_bs, int gh_accumulate, float* g_x0, long long B,
                          int F, int H, int D, int L, float* scratch, cudaStream_t st) {
    using namespace tc;
    const int FP = cin_tc_fp(F);
    DFM_REQUIRE(FP > 0 && L <= 256, DFM_ERR_UNSUPPORTED, "cin tcgen05 backward: F=%d / L=%d outside the instantiated tiles", F, L);
    const int HT = 256 / FP, NT = HT * FP;
    const int Lp = (L + 31) & ~31, Hp = (H + HT - 1) / HT * HT;
    const long long M = B * D;
    float* gT = scratch;
    float* wt = scratch + (((size_t)M * Lp + 31) & ~(size_t)31);
    const size_t tile_smem = (size_t)L * (D + 1) * 4;
    DFM_REQUIRE(tile_smem <= 200 * 1024, DFM_ERR_UNSUPPORTED, "cin tcgen05 backward: L*D too large for the transpose tile");
    if (g_pre) {      // nullptr: g_pre^T is already in the scratch (cin_gpre_fused)
        if (tile_smem > 48 * 1024)
            DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_gpre_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem));
        cin_gpre_transpose_kernel<<<(unsigned)B, 256, tile_smem, st>>>(g_pre, B, L, D, Lp, gT);
    }
    long long pb = ceil_div((long long)Hp * FP * Lp, 256);
    if (pb > 4LL * sm_count()) pb = 4LL * sm_count();
    cin_wt_pad_kernel<<<(unsigned)pb, 256, 0, st>>>(w, L, H, F, FP, Hp, Lp, wt);
    DFM_CHECK_LAUNCH();
    CUtensorMap amap, bmap;
    int rc = make_tmap_2d(&amap, gT, M, Lp, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&bmap, wt, (long long)Hp * FP, Lp, NT);
    if (rc) return rc;
    BwdDataArgs a;
    a.x0 = x0; a.x_bs = x_bs; a.hid = hid; a.h_bs = h_bs; a.g_hid = g_hid; a.gh_bs = gh_bs; a.g_x0 = g_x0;
    a.M = M; a.F = F; a.H = H; a.D = D; a.Lp = Lp; a.n_tiles = Hp / HT; a.gh_accumulate = gh_accumulate;
    const size_t a_bytes = (size_t)(Lp / 32) * 128 * 128, b_stage = ((size_t)NT * 128 + 1023) & ~(size_t)1023;
    int nstage = BW_NSTAGE;
    while (nstage > 1 && a_bytes + nstage * b_stage + 256 + 1024 > 227 * 1024) --nstage;
    a.nstage = nstage;
    const size_t smem = a_bytes + nstage * b_stage + 256 + 1024;
    DFM_REQUIRE(smem <= 227 * 1024, DFM_ERR_UNSUPPORTED, "cin tcgen05 backward: %zu B shared memory", smem);
    long long grid = ceil_div(M, 128);
    if (grid > sm_count()) grid = sm_count();
    switch (FP) {
        case 16: return launch_bwd_data<16>(a, amap, bmap, smem, (int)grid, st);
        case 24: return launch_bwd_data<24>(a, amap, bmap, smem, (int)grid, st);
        case 32: return launch_bwd_data<32>(a, amap, bmap, smem, (int)grid, st);
        case 40: return launch_bwd_data<40>(a, amap, bmap, smem, (int)grid, st);
        case 48: return launch_bwd_data<48>(a, amap, bmap, smem, (int)grid, st);
        default: return launch_bwd_data<64>(a, amap, bmap, smem, (int)grid, st);
    }
}

}  // namespace dfm

// =====================================================================================================
// Weight gradient on tcgen05, transposed form:  dW^T[k'][l] = sum_r z[r][k'] * g_pre[r][l],
//   k' = h*FP + f (128 rows per CTA tile = UMMA M), N = Lp, reduction over the rows r = (b,d).
// The A operand z[r][k'] = hidden[r][h] * x0[r][f] is synthesised per 32-row block: the hidden / x0
// values of the block are staged in (padded) shared memory, 8 producer warps (row k' x half block)
// multiply them and write the tile straight into tensor memory (tcgen05.st, TS-mode MMA); the B operand
// g_pre as (Lp, M) row-major arrives by TMA.  One CTA = one k' tile x one slice of rows; the partial
// dW^T of every slice is reduced in slice order afterwards (deterministic).
namespace dfm {
namespace tc {

constexpr int DW_NSTAGE = 4;
constexpr int DW_RB = 64;                       // reduction rows per k-block (stage): 2 TMA boxes of 32, 8 MMAs of K = 8
constexpr int DW_PD = 2;                        // prefetch distance of the producers' global loads, in k-blocks
constexpr int DW_NQ = 5;                        // staged quads per producer thread and k-block: (nh + F) * 16 <= 5 * 256
constexpr int DW_PRODUCERS = 256;
constexpr int DW_THREADS = DW_PRODUCERS + 64;   // + MMA warp + TMA warp
constexpr int SLAB = DW_RB + 4;                 // floats per staged channel row: 64 r + 4 pad (bank spread)

__device__ __forceinline__ void tmem_st16(uint32_t addr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct DwTcArgs {
    const float* x0; long long x_bs;
    const float* hid; long long h_bs;
    float* part;                 // (n_slices, Kt, Lp)   Kt = n_ktiles * 128
    long long M, slice_rows;
    int F, FP, H, D, Lp, Kt;
    int nstage;                  // B / A ring depth (<= DW_NSTAGE; fewer when Lp = 256 fills the shared memory)
};

__global__ void __launch_bounds__(DW_THREADS, 1)
cin_tc_dw_kernel(const __grid_constant__ DwTcArgs a, const __grid_constant__ CUtensorMap gmap) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int ktile = blockIdx.x, slice = blockIdx.y;
    const int b_box = (a.Lp * 128 + 1023) & ~1023;          // one TMA box: 32 r x Lp rows
    const int b_stage = (DW_RB / 32) * b_box;
    const int h_lo = (ktile * 128) / a.FP;
    const int h_hi = (ktile * 128 + 127) / a.FP;
    const int nh = h_hi - h_lo + 1;
    const int nch = nh + a.F;                               // staged channels: nh hidden rows, then F x0 rows
    unsigned char* sB = smem;
    float* slab = reinterpret_cast<float*>(smem + (size_t)a.nstage * b_stage);           // [2][nch][SLAB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(slab + (size_t)2 * nch * SLAB);
    uint64_t* full = bars;                   // [DW_NSTAGE] 8 producer warps + TMA
    uint64_t* empty = bars + DW_NSTAGE;      // [DW_NSTAGE] MMA commit
    uint64_t* acc_full = bars + 2 * DW_NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DW_NSTAGE; ++s) { mbar_init(full + s, 9); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&gmap)) : "memory");
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_col0 = (uint32_t)a.Lp;                 // A ring behind the accumulator
    const long long r_lo = (long long)slice * a.slice_rows;
    const long long r_hi = (r_lo + a.slice_rows < a.M) ? r_lo + a.slice_rows : a.M;
    const int n_kb = (int)((r_hi - r_lo + DW_RB - 1) / DW_RB);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.Lp >> 3) << 17) | ((128u >> 4) << 24);

    if (warp < 8) {
        // ---------------------------------------------------------------- producers
        const int t = threadIdx.x, row = t & 127, half = t >> 7;
        const int kp = ktile * 128 + row;
        const int h = kp / a.FP, f = kp - h * a.FP;
        const bool row_ok = h < a.H && f < a.F;
        const int hs = (h - h_lo) * SLAB + half * (DW_RB / 2), xs = (nh + f) * SLAB + half * (DW_RB / 2);
        // slab staging assignment: quads of 4 consecutive r of one channel.  Loads are issued DW_PD k-blocks ahead of
        // their use (registers pre[j][u]): one k-block is only 4 MMAs = 256 tensor cycles, a global / L2 load takes
        // ~1000, so a prefetch distance of one k-block left the tensor pipe idle 85 % of the time.  Addresses advance
        // incrementally (b, d += 32 rows per k-block): no division in the loop.
        constexpr int QPC = DW_RB / 4;       // quads per channel
        const int n_quads = nch * QPC;
        const float* qbase[DW_NQ];  // channel base pointer (hidden row h_lo + c, or x0 row c - nh); null = all-zero channel
        long long qbs[DW_NQ];       // batch stride of that tensor
        long long qb[DW_NQ];        // sample index of the NEXT load of this quad
        int qd[DW_NQ];              // d of the next load
        long long qr[DW_NQ];        // global row r of the next load
#pragma unroll
        for (int u = 0; u < DW_NQ; ++u) {
            const int idx = t + u * DW_PRODUCERS;
            const int c = idx / QPC, q = idx % QPC;
            const bool ok = idx < n_quads && !(c < nh && h_lo + c >= a.H);
            qbase[u] = !ok ? nullptr : (c < nh ? a.hid + (size_t)(h_lo + c) * a.D : a.x0 + (size_t)(c - nh) * a.D);
            qbs[u] = c < nh ? a.h_bs : a.x_bs;
            qr[u] = r_lo + 4 * q;
            qb[u] = qr[u] / a.D;
            qd[u] = (int)(qr[u] - qb[u] * a.D);
        }
        auto next_quad = [&](int u) -> float4 {          // load the quad at (qb, qd), then advance it by one k-block (32 rows)
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (qbase[u] && qr[u] < r_hi) v = __ldg(reinterpret_cast<const float4*>(qbase[u] + qb[u] * qbs[u] + qd[u]));
            qr[u] += DW_RB;
            qd[u] += DW_RB;
            while (qd[u] >= a.D) { qd[u] -= a.D; ++qb[u]; }
            return v;
        };
        float4 pre[DW_PD][DW_NQ];
#pragma unroll
        for (int j = 0; j < DW_PD; ++j)
#pragma unroll
            for (int u = 0; u < DW_NQ; ++u) pre[j][u] = next_quad(u);   // k-blocks beyond the slice read as zeros (r >= r_hi)
        uint32_t s = 0, ph = 0;
        for (int kb0 = 0; kb0 < n_kb; kb0 += DW_PD) {
#pragma unroll
          for (int j = 0; j < DW_PD; ++j) {
            const int kb = kb0 + j;
            if (kb >= n_kb) break;                           // uniform over the 256 producers
            float* sl = slab + (size_t)(kb & 1) * nch * SLAB;
#pragma unroll
            for (int u = 0; u < DW_NQ; ++u) {
                const int idx = t + u * DW_PRODUCERS;
                if (idx < n_quads) *reinterpret_cast<float4*>(sl + (idx / QPC) * SLAB + (idx % QPC) * 4) = pre[j][u];
            }
            producer_bar();                                  // slab kb complete (slab kb-1 was consumed before its own bar)
#pragma unroll
            for (int u = 0; u < DW_NQ; ++u) pre[j][u] = next_quad(u);  // k-block kb + DW_PD
            float z[DW_RB / 2];
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < DW_RB / 8; ++q) {
                    const float4 hv = *reinterpret_cast<const float4*>(sl + hs + 4 * q);
                    const float4 xv = *reinterpret_cast<const float4*>(sl + xs + 4 * q);
                    z[4 * q] = hv.x * xv.x; z[4 * q + 1] = hv.y * xv.y; z[4 * q + 2] = hv.z * xv.z; z[4 * q + 3] = hv.w * xv.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < DW_RB / 2; ++i) z[i] = 0.f;
            }
            mbar_wait(empty + s, ph ^ 1u);
            tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + a_col0 + s * DW_RB + half * (DW_RB / 2), z);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + s);
            if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
          }
        }
        // ---------------------------------------------------------------- epilogue (warps 0-3: lane = k' row)
        if (warp < 4) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
            float* out = a.part + ((size_t)slice * a.Kt + (size_t)ktile * 128 + (warp * 32 + lane)) * a.Lp;
            for (int c0 = 0; c0 < a.Lp; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(out + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            tc_fence_before();
        }
    } else if (warp == 9 && lane == 0) {
        // ---------------------------------------------------------------- TMA: g_pre (Lp, M) box 32 r x Lp rows
        uint32_t s = 0, ph = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(empty + s, ph ^ 1u);
            mbar_arrive_expect_tx(full + s, (uint32_t)(DW_RB / 32) * (uint32_t)a.Lp * 128u);
#pragma unroll
            for (int bx = 0; bx < DW_RB / 32; ++bx)       // rows beyond M are zero-filled by the TMA unit
                tma_load_2d(sB + (size_t)s * b_stage + (size_t)bx * b_box, &gmap, (int)(r_lo + (long long)kb * DW_RB + 32 * bx), 0, full + s);
            if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 8 && lane == 0) {
        // ---------------------------------------------------------------- MMA issuer
        uint32_t s = 0, ph = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(full + s, ph);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < DW_RB / 8; ++k) {
                const uint64_t db = make_desc(smem_u32(sB + (size_t)s * b_stage + (size_t)(k >> 2) * b_box));
                umma_tf32_ts(tmem_base, tmem_base + a_col0 + s * DW_RB + k * 8, db + (uint64_t)((k & 3) * 2), idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit(empty + s);
            if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
        }
        umma_commit(acc_full);
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// g_pre (B, L, D) -> (Lp, M) row-major (column r = b*D + d), zero padded rows l >= L
__global__ void cin_gpre_lm_kernel(const float* __restrict__ gp, long long B, int L, int D, int Lp, long long Mld,
                                   float* __restrict__ out) {
    const long long n = (long long)Lp * B * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i % (B * D);
        const int l = (int)(i / (B * D));
        const long long b = r / D;
        const int d = (int)(r - b * D);
        out[(size_t)l * Mld + r] = l < L ? __ldg(gp + (b * L + l) * D + d) : 0.f;
    }
}

// gW[l][h*F+f] = sum_slices part[s][h*FP+f][l]
__global__ void cin_dw_reduce_kernel(const float* __restrict__ part, int n_slices, int Kt, int Lp, int L, int H, int F,
                                     int FP, float* __restrict__ gw) {
    const long long n = (long long)L * H * F;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = (int)(i / (H * F)), k = (int)(i - (long long)l * H * F);
    const int h = k / F, f = k - h * F;
    const size_t src = (size_t)(h * FP + f) * Lp + l;
    float acc = 0.f;
    for (int s = 0; s < n_slices; ++s) acc += part[(size_t)s * Kt * Lp + src];
    gw[i] = acc;
}

// gb[l] = sum_r g_pre[b][l][d] : block partials then fixed-order final sum
__global__ void __launch_bounds__(256)
cin_db_partial_kernel(const float* __restrict__ gp, long long B, int L, int D, float* __restrict__ part) {
    // block y = l, block x = slice of samples; 256 threads stride over (b, d)
    __shared__ float red[8];
    const int l = blockIdx.y;
    const long long per = (B + gridDim.x - 1) / gridDim.x, b0 = (long long)blockIdx.x * per;
    const long long b1 = b0 + per < B ? b0 + per : B;
    float acc = 0.f;
    for (long long i = b0 * D + threadIdx.x; i < b1 * D; i += 256) {
        const long long b = i / D;
        acc += __ldg(gp + (b * L + l) * D + (i - b * D));
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; part[(size_t)blockIdx.x * L + l] = s; }
}
// the same partial sums from the (Lp, M) layout: row l is contiguous
__global__ void __launch_bounds__(256)
cin_db_rows_kernel(const float* __restrict__ gLM, long long M, int L, float* __restrict__ part) {
    __shared__ float red[8];
    const int l = blockIdx.y;
    const long long per = (M + gridDim.x - 1) / gridDim.x, r0 = (long long)blockIdx.x * per;
    const long long r1 = r0 + per < M ? r0 + per : M;
    const float* row = gLM + (size_t)l * M;
    float acc = 0.f;
    for (long long i = r0 + threadIdx.x; i < r1; i += 256) acc += __ldg(row + i);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float s = 0.f; for (int w = 0; w < 8; ++w) s += red[w]; part[(size_t)blockIdx.x * L + l] = s; }
}

// g_pre = (act > 0) * ([l < direct] g_out[b, col_off + l] + [l in next range] g_hnext[b, l - next_off, d])  (cin.py:91-102
// backward) written ONCE, straight into the two layouts the tensor-core kernels read: g_pre^T (B*D, Lp) for the
// data-gradient GEMM and (Lp, B*D) for the weight-gradient GEMM (both zero padded to Lp) -- the (B, L, D) copy, its two
// re-layout passes and their re-reads never happen.  One block per group of `spb` samples, through a shared-memory tile.
// index split without integer division when the divisor is a power of two (lg >= 0)
__device__ __forceinline__ void split(int i, int n, int lg, int& q, int& r) {
    if (lg >= 0) { q = i >> lg; r = i & (n - 1); } else { q = i / n; r = i - q * n; }
}

__global__ void __launch_bounds__(256)
cin_gpre_fused_kernel(const float* __restrict__ act, const float* __restrict__ g_out, const float* __restrict__ g_hnext,
                      long long B, int L, int D, int Lp, int direct, int out_dim, int col_off, int next_off, int next_n,
                      int spb, int lgD4, int lgLp4, float* __restrict__ gT, float* __restrict__ gLM) {
    extern __shared__ float tile[];                      // [spb][L][D + 1]
    const long long b0 = (long long)blockIdx.x * spb;
    const int ns = (int)((B - b0 < spb) ? B - b0 : spb);
    const int LD = L * D, D1 = D + 1, D4 = D >> 2, Lp4 = Lp >> 2;
    const long long M = B * D;
    // 1. g_pre of ns samples -> tile; 128-bit reads of the activation (d is contiguous)
    for (int s = 0; s < ns; ++s) {
        const long long b = b0 + s;
        const float4* a4 = reinterpret_cast<const float4*>(act + b * LD);
        float* ts = tile + (size_t)s * L * D1;
        for (int i = threadIdx.x; i < (LD >> 2); i += 256) {
            int l, c;
            split(i, D4, lgD4, l, c);
            const float4 a = __ldcs(a4 + i);
            float gd = 0.f;
            if (l < direct) gd = __ldg(g_out + b * out_dim + col_off + l);
            float4 gh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g_hnext && l >= next_off && l < next_off + next_n)
                gh = __ldcs(reinterpret_cast<const float4*>(g_hnext + (b * next_n + (l - next_off)) * D) + c);
            float* t = ts + l * D1 + (c << 2);
            t[0] = a.x > 0.f ? gd + gh.x : 0.f; t[1] = a.y > 0.f ? gd + gh.y : 0.f;
            t[2] = a.z > 0.f ? gd + gh.z : 0.f; t[3] = a.w > 0.f ? gd + gh.w : 0.f;
        }
    }
    __syncthreads();
    // 2. g_pre^T (B*D, Lp): rows (b, d), 128-bit stores along l
    for (int i = threadIdx.x; i < ns * D * Lp4; i += 256) {
        int sd, l4;
        split(i, Lp4, lgLp4, sd, l4);
        int s, d;
        split(sd, D, lgD4 >= 0 ? lgD4 + 2 : -1, s, d);
        const int l = l4 << 2;
        const float* t = tile + ((size_t)s * L + l) * D1 + d;
        float4 v;
        v.x = l < L ? t[0] : 0.f; v.y = l + 1 < L ? t[D1] : 0.f; v.z = l + 2 < L ? t[2 * D1] : 0.f; v.w = l + 3 < L ? t[3 * D1] : 0.f;
        __stcs(reinterpret_cast<float4*>(gT + ((b0 * D + sd) * Lp)) + l4, v);
    }
    // 3. (Lp, M): for every l the ns * D values of this sample group are contiguous, 128-bit stores along (s, d)
    const int span4 = ns * D4;
    for (int i = threadIdx.x; i < Lp * span4; i += 256) {
        const int l = i / span4, c = i - l * span4;
        int s, c4;
        split(c, D4, lgD4, s, c4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < L) { const float* t = tile + ((size_t)s * L + l) * D1 + (c4 << 2); v = make_float4(t[0], t[1], t[2], t[3]); }
        __stcs(reinterpret_cast<float4*>(gLM + (size_t)l * M + b0 * D) + c, v);
    }
}

__global__ void cin_db_final_kernel(const float* __restrict__ part, int n, int L, float* __restrict__ gb) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    float acc = 0.f;
    for (int s = 0; s < n; ++s) acc += part[(size_t)s * L + l];
    gb[l] = acc;
}

}  // namespace tc

static int dw_tc_slices(long long M, int n_ktiles) {
    // one CTA per SM (shared-memory bound): k-tiles x slices must not spill into a third, nearly empty wave
    // (20 k-tiles x ceil(296 / 20) = 300 CTAs ran 3 rounds on 148 SMs: tensor pipe 30 % active but 20 % elapsed)
    long long want = (2LL * sm_count()) / n_ktiles;
    long long max_s = ceil_div(M, 4096);
    if (want > max_s) want = max_s;
    if (want < 1) want = 1;
    return (int)want;
}

size_t cin_tc_dw_scratch_floats(long long B, int F, int D, int Hmax, int Lmax) {
    const int FP = cin_tc_fp(F);
    if (!FP) return 0;
    const int Lp = (Lmax + 31) & ~31;
    const int n_ktiles = (int)ceil_div((long long)Hmax * FP, 128);
    const long long M = B * D;
    const int ns = dw_tc_slices(M, n_ktiles);
    return (size_t)Lp * M + (size_t)ns * n_ktiles * 128 * Lp + 64 * (size_t)Lmax + 256;
}

// Can both tensor-core backward kernels take this layer (same conditions as their own checks)?
bool cin_tc_bwd_supported(long long B, int F, int H, int D, int L) {
    const int FP = cin_tc_fp(F);
    if (!(FP > 0 && L <= 256 && D % 4 == 0)) return false;
    const int HT = 256 / FP, NT = HT * FP, Lp = (L + 31) & ~31;
    const size_t a_bytes = (size_t)(Lp / 32) * 128 * 128, b_stage = ((size_t)NT * 128 + 1023) & ~(size_t)1023;
    if (a_bytes + b_stage + 256 + 1024 > 227 * 1024) return false;
    const int nh_max = 127 / FP + 2;
    // dW kernel: at least 2 stages of (DW_RB / 32) boxes + the double-buffered slab, and the staging assignment must fit
    const size_t smem = (size_t)2 * (tc::DW_RB / 32) * ((Lp * 128 + 1023) & ~1023) + (size_t)2 * (nh_max + F) * tc::SLAB * 4 + 256 + 1024;
    if ((nh_max + F) * (tc::DW_RB / 4) > tc::DW_NQ * tc::DW_PRODUCERS) return false;
    return smem <= 227 * 1024 && (size_t)L * (D + 1) * 4 <= 64 * 1024;
}

// g_pre of one layer straight into the scratch layouts of cin_layer_bwd_data_tc (gT) and cin_layer_dw_tc (gLM)
int cin_gpre_fused(const float* act, const float* g_out, const float* g_hnext, long long B, int L, int D, int direct,
                   int out_dim, int col_off, int next_off, int next_n, float* bwd_scratch, float* dw_scratch, cudaStream_t st) {
    using namespace tc;
    const int Lp = (L + 31) & ~31;
    const size_t per = (size_t)L * (D + 1) * 4;
    int spb = (int)((64 * 1024) / per);
    if (spb < 1) spb = 1;
    if (spb > 8) spb = 8;
    const size_t smem = per * spb;
    if (smem > 48 * 1024)
        DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_gpre_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto lg = [](int n) { int b = 0; while ((1 << b) < n) ++b; return (1 << b) == n ? b : -1; };
    cin_gpre_fused_kernel<<<(unsigned)ceil_div(B, spb), 256, smem, st>>>(act, g_out, g_hnext, B, L, D, Lp, direct, out_dim, col_off,
                                                                          next_off, next_n, spb, lg(D / 4), lg(Lp / 4), bwd_scratch,
                                                                          dw_scratch);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

// dW and db of one CIN layer on tcgen05
int cin_layer_dw_tc(const float* g_pre, const float* x0, long long x_bs, const float* hid, long long h_bs, float* gw,
                    float* gb, long long B, int F, int H, int D, int L, float* scratch, cudaStream_t st) {
    using namespace tc;
    const int FP = cin_tc_fp(F);
    DFM_REQUIRE(FP > 0 && L <= 256 && D % 4 == 0, DFM_ERR_UNSUPPORTED, "cin tcgen05 dW: F=%d L=%d D=%d outside the instantiated tiles", F, L, D);
    const int Lp = (L + 31) & ~31;
    const long long M = B * D;
    const int n_ktiles = (int)ceil_div((long long)H * FP, 128), Kt = n_ktiles * 128;
    const int ns = dw_tc_slices(M, n_ktiles);
    const long long slice_rows = ceil_div(ceil_div(M, ns), DW_RB) * DW_RB;
    const int real_slices = (int)ceil_div(M, slice_rows);
    float* gLM = scratch;
    float* part = scratch + (((size_t)Lp * M + 63) & ~(size_t)63);
    float* dbp = part + (size_t)ns * Kt * Lp;
    long long gb_ = ceil_div((long long)Lp * M, 256);
    if (gb_ > 16LL * sm_count()) gb_ = 16LL * sm_count();
    if (g_pre) cin_gpre_lm_kernel<<<(unsigned)gb_, 256, 0, st>>>(g_pre, B, L, D, Lp, M, gLM);   // nullptr: already there
    DFM_CHECK_LAUNCH();
    CUtensorMap gmap;
    int rc = make_tmap_2d(&gmap, gLM, Lp, M, Lp);
    if (rc) return rc;
    DwTcArgs a;
    a.x0 = x0; a.x_bs = x_bs; a.hid = hid; a.h_bs = h_bs; a.part = part; a.M = M; a.slice_rows = slice_rows;
    a.F = F; a.FP = FP; a.H = H; a.D = D; a.Lp = Lp; a.Kt = Kt;
    const int nh_max = 127 / FP + 2;
    DFM_REQUIRE((nh_max + F) * (DW_RB / 4) <= DW_NQ * DW_PRODUCERS, DFM_ERR_UNSUPPORTED, "cin tcgen05 dW: %d staged channels", nh_max + F);
    const size_t b_stage = (size_t)(DW_RB / 32) * ((Lp * 128 + 1023) & ~1023);
    const size_t fixed = (size_t)2 * (nh_max + F) * SLAB * 4 + 256 + 1024;
    int nstage = DW_NSTAGE;
    while (nstage > 2 && ((size_t)nstage * b_stage + fixed > 227 * 1024 || Lp + nstage * DW_RB > 512)) --nstage;
    a.nstage = nstage;
    const size_t smem = (size_t)nstage * b_stage + fixed;
    DFM_REQUIRE(smem <= 227 * 1024, DFM_ERR_UNSUPPORTED, "cin tcgen05 dW: %zu B shared memory", smem);
    DFM_CHECK_CUDA(cudaFuncSetAttribute(cin_tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cin_tc_dw_kernel<<<dim3(n_ktiles, real_slices), DW_THREADS, smem, st>>>(a, gmap);
    cin_dw_reduce_kernel<<<(unsigned)ceil_div((long long)L * H * F, 256), 256, 0, st>>>(part, real_slices, Kt, Lp, L, H, F, FP, gw);
    const int dbs = 64;
    if (g_pre) cin_db_partial_kernel<<<dim3(dbs, L), 256, 0, st>>>(g_pre, B, L, D, dbp);
    else cin_db_rows_kernel<<<dim3(dbs, L), 256, 0, st>>>(gLM, M, L, dbp);
    cin_db_final_kernel<<<(unsigned)ceil_div(L, 128), 128, 0, st>>>(dbp, dbs, L, gb);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // namespace dfm
