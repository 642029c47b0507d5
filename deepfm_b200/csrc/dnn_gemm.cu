// fp32-accurate GEMMs of the DNN tower on the 5th-generation tensor cores (reference:
// deepfm/models/layers/dnn.py:45-59 -- nn.Linear forward and its two autograd products, ATen mm / addmm).
//
// One persistent, warp-specialised tcgen05 kernel computes D[M][N] = sum_k A(m,k) * B(n,k) in "3xTF32": every fp32
// operand x is split into x_hi (its top 19 bits, a valid tf32 number) and x_lo = x - x_hi (exact in fp32), and
//     D += A_hi * B_hi + A_hi * B_lo + A_lo * B_hi          (fp32 accumulation in tensor memory)
// so the dropped term is ~2^-22 relative: the accuracy class of an fp32 SIMT sgemm at 3 tensor-core passes.
// Operand layouts (no transposed copies anywhere):
//     mode 0  NT   A [M][K]   B [N][K]     Y  = X  W^T (+ bias)    both K-major
//     mode 1  NN   A [M][K]   B [K][N]     dX = dY W               A K-major, B MN-major
//     mode 2  TN   A [K][M]   B [K][N]     dW = dY^T X             both MN-major, split-K with a fixed-order reduce
// Pipeline per k-block of 32:  warp 0 issues the TMA loads (128-byte-swizzled boxes, mbarrier complete_tx);  warps
// 2-5 split the landed tiles IN PLACE (raw -> hi), write B_lo to a second shared-memory tile and A_lo straight
// into tensor memory (tcgen05.st: the A_lo * B_hi product runs in the TS form, which halves its shared-memory
// reads);  one elected thread of warp 1 issues 12 `tcgen05.mma.cta_group::1.kind::tf32` (M = 128, N = tile, K = 8)
// and recycles the stage with tcgen05.commit;  warps 6-9 are the PROMOTERS: the tensor core's fp32 accumulation
// rounds toward zero, so a long chain of accumulating MMAs drifts (measured 2e-5 relative at K = 2496); the MMAs
// therefore accumulate only CH k-blocks (K = 128) into one of two TMEM accumulators, and the promoter warps add
// each finished chunk (tcgen05.ld) into round-to-nearest fp32 master accumulators held in registers while the
// next chunk's MMAs fill the other TMEM buffer; the masters go to global memory at the end of the tile.
// MN-major 32-bit operands must use the 32-byte-atom 128-byte swizzle (TMA SWIZZLE_128B_ATOM_32B ==
// UMMA SWIZZLE_128B_BASE32B: 32-byte chunks XOR-ed with the row index mod 4).
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace dfm {
namespace g3 {

constexpr int TM = 128;          // rows of D per tile (UMMA M)
constexpr int KB = 32;           // reduction elements per stage: 32 tf32 = one 128-byte swizzle row
constexpr int MAXST = 6;
constexpr int CH = 4;            // k-blocks per TMEM accumulation chunk (K = 128: 48 accumulating MMAs)
constexpr int TN_MAX = 128;      // N tile: the promoters keep one master accumulator per column in registers
constexpr int THREADS = 320;     // warp 0: TMA, warp 1: MMA, warps 2-5: split, warps 6-9: epilogue
constexpr int A_BYTES = TM * 128;

struct Args {
    float* D; long long ldd;             // output (or split-K partial base), row stride in floats
    const float* bias;                   // (N) or null
    long long M, N, K;
    int tn;                              // N tile: multiple of 32, <= TN_MAX
    int n_mt, n_nt, n_split;
    long long k_per_split;               // multiple of KB
    long long split_stride;              // floats between the partial outputs of consecutive splits
    int a_mn, b_mn;                      // 1: the operand is MN-major ([K][MN] in memory)
    int nstage;
    int mask_hi;                         // 1: write x_hi back over the raw tile (0, default: kind::tf32 ignores the low 13
                                         // mantissa bits of its operands -- measured: identical results -- so raw IS hi)
    int b_lo_tma;                        // 1: B_lo was precomputed in global memory (weights) and arrives by TMA (mapBlo)
    int ts_all;                          // 1: A_hi goes to tensor memory too, all three products run in the TS form (the
                                         // tensor core then reads only B tiles from shared memory: 48 KB instead of 80 KB
                                         // per k-block -- shared-memory bandwidth, not the tensor pipe, bounds this kernel)
    int tma_store;                       // 1: the epilogue stages 32 x 32 boxes in shared memory and stores them by TMA
                                         // (coalesced 128-byte rows; a thread-per-row float4 store touches 32 half-written
                                         // sectors per instruction and made the LSU the bound of every small-K product)
    int transpose_out;                   // 1: the tile is stored transposed (D[n][m]): the swapped weight-gradient product
    uint32_t tmem_cols;
};
constexpr int EPI_BYTES = 4 * 2 * 4096; // 4 promoter warps x 2 buffers x (32 rows x 128 B)

// MN-major 32-bit operand: cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B (canonical layout ((8,n),(4,k)):((1,LBO),(8,SBO))
// in 16-byte units): 32 MN elements are contiguous (128 B), 4 k-rows of 128 B form one 512-byte swizzle atom (32-byte
// chunks XOR-ed with the row index mod 4), the next 4 k-rows follow SBO = 512 B later, the next 32 MN elements start
// LBO = 32 k-rows * 128 B = 4096 B further (one TMA box {32 mn, 32 k}, SWIZZLE_128B_ATOM_32B, per MN atom).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;       // leading byte offset: next MN atom
    d |= (uint64_t)(512 >> 4) << 32;        // stride byte offset: next group of 4 k
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(THREADS, 1)
gemm3_kernel(const __grid_constant__ Args a, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapBlo, const __grid_constant__ CUtensorMap mapD) {
    using namespace tc;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int b_bytes = a.tn * 128;
    const int stage_bytes = A_BYTES + 2 * b_bytes;          // [A raw -> hi][B raw -> hi][B lo]
    unsigned char* epi = smem + (size_t)a.nstage * stage_bytes;                    // 1024-byte aligned (stages are multiples of 4 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi + (a.tma_store ? EPI_BYTES : 0));
    uint64_t* full_raw = bars;                 // TMA landed
    uint64_t* full_split = bars + MAXST;       // hi / lo tiles ready
    uint64_t* empty = bars + 2 * MAXST;        // MMAs of the stage retired
    uint64_t* acc_full = bars + 3 * MAXST;     // [2]
    uint64_t* acc_empty = acc_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < MAXST; ++s) { mbar_init(full_raw + s, 1); mbar_init(full_split + s, 4); mbar_init(empty + s, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t alo_col0 = (uint32_t)(2 * a.tn);          // TMEM columns of the A ring: per stage [A_lo 32][A_hi 32 if ts_all]
    const uint32_t a_cols = a.ts_all ? 64u : 32u;
    const long long n_units = (long long)a.n_mt * a.n_nt * a.n_split;

    auto unit_of = [&](long long u, int& mt, int& nt, int& sp) {
        nt = (int)(u % a.n_nt);
        const long long rest = u / a.n_nt;
        mt = (int)(rest % a.n_mt);
        sp = (int)(rest / a.n_mt);
    };
    auto kblocks = [&](int sp, long long& k0) {
        k0 = (long long)sp * a.k_per_split;
        long long k1 = k0 + a.k_per_split;
        if (k1 > a.K) k1 = a.K;
        return (int)((k1 - k0 + KB - 1) / KB);
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            const uint32_t tx = (uint32_t)(A_BYTES + b_bytes * (a.b_lo_tma ? 2 : 1));
            for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
                int mt, nt, sp;
                unit_of(u, mt, nt, sp);
                long long k0;
                const int nkb = kblocks(sp, k0);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty + s, ph ^ 1u);
                    mbar_arrive_expect_tx(full_raw + s, tx);
                    unsigned char* sa = smem + (size_t)s * stage_bytes;
                    unsigned char* sb = sa + A_BYTES;
                    const int kk = (int)(k0 + (long long)kb * KB);
                    if (!a.a_mn) tma_load_2d(sa, &mapA, kk, mt * TM, full_raw + s);                   // box {32 k, 128 rows}
                    else for (int j = 0; j < TM / 32; ++j) tma_load_2d(sa + j * 4096, &mapA, mt * TM + j * 32, kk, full_raw + s);
                    if (!a.b_mn) tma_load_2d(sb, &mapB, kk, nt * a.tn, full_raw + s);                 // box {32 k, tn rows}
                    else for (int j = 0; j < a.tn / 32; ++j) tma_load_2d(sb + j * 4096, &mapB, nt * a.tn + j * 32, kk, full_raw + s);
                    if (a.b_lo_tma) {
                        unsigned char* sl = sb + b_bytes;
                        if (!a.b_mn) tma_load_2d(sl, &mapBlo, kk, nt * a.tn, full_raw + s);
                        else for (int j = 0; j < a.tn / 32; ++j) tma_load_2d(sl + j * 4096, &mapBlo, nt * a.tn + j * 32, kk, full_raw + s);
                    }
                    if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.tn >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc_ss = idesc | ((uint32_t)a.a_mn << 15) | ((uint32_t)a.b_mn << 16);
            const uint32_t idesc_ts = idesc | ((uint32_t)a.b_mn << 16);
            const uint64_t adv_a = a.a_mn ? (1024 >> 4) : (32 >> 4);     // descriptor step per K = 8
            const uint64_t adv_b = a.b_mn ? (1024 >> 4) : (32 >> 4);
            long long it = 0;                                        // accumulation chunks issued so far (both TMEM buffers)
            for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
                int mt, nt, sp;
                unit_of(u, mt, nt, sp);
                long long k0;
                const int nkb = kblocks(sp, k0);
                uint32_t tacc = 0, acc = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (kb % CH == 0) {                              // a new chunk starts in the other accumulator
                        acc = (uint32_t)(it & 1);
                        mbar_wait(acc_empty + acc, (uint32_t)((it >> 1) & 1) ^ 1u);     // the promoters drained it
                        tc_fence_after();
                        tacc = tmem_base + acc * (uint32_t)a.tn;
                    }
                    mbar_wait(full_raw + s, ph);
                    mbar_wait(full_split + s, ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t da = a.a_mn ? make_desc_mn(sa) : make_desc(sa);
                    const uint64_t db = a.b_mn ? make_desc_mn(sa + A_BYTES) : make_desc(sa + A_BYTES);
                    const uint64_t dbl = a.b_mn ? make_desc_mn(sa + A_BYTES + b_bytes) : make_desc(sa + A_BYTES + b_bytes);
                    const uint32_t talo = tmem_base + alo_col0 + s * a_cols;
                    if (a.ts_all) {
                        const uint32_t tahi = talo + 32;
#pragma unroll
                        for (int k = 0; k < KB / 8; ++k) {
                            umma_tf32_ts(tacc, tahi + k * 8, db + adv_b * k, idesc_ts, ((kb % CH) | k) ? 1u : 0u);  // hi * hi
                            umma_tf32_ts(tacc, tahi + k * 8, dbl + adv_b * k, idesc_ts, 1u);                        // hi * lo
                            umma_tf32_ts(tacc, talo + k * 8, db + adv_b * k, idesc_ts, 1u);                         // lo * hi
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < KB / 8; ++k) {
                            umma_tf32(tacc, da + adv_a * k, db + adv_b * k, idesc_ss, ((kb % CH) | k) ? 1u : 0u);   // hi * hi
                            umma_tf32(tacc, da + adv_a * k, dbl + adv_b * k, idesc_ss, 1u);                         // hi * lo
                            umma_tf32_ts(tacc, talo + k * 8, db + adv_b * k, idesc_ts, 1u);                         // lo * hi
                        }
                    }
                    umma_commit(empty + s);
                    if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
                    if (kb % CH == CH - 1 || kb == nkb - 1) { umma_commit(acc_full + acc); ++it; }
                }
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------ splitters: raw -> hi (in place), lo
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                    // A row (= TMEM lane)
        const int tid = threadIdx.x - 64;
        uint32_t s = 0, ph = 0;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            int mt, nt, sp;
            unit_of(u, mt, nt, sp);
            long long k0;
            const int nkb = kblocks(sp, k0);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(full_raw + s, ph);
                unsigned char* sa = smem + (size_t)s * stage_bytes;
                float lo[32], hi[32];                    // hi = the raw value: kind::tf32 ignores the low 13 mantissa bits
                if (!a.a_mn) {                           // row r: 8 chunks of 4 k, chunk c at ((c ^ (r & 7)) << 4)
                    unsigned char* rowp = sa + r * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4* p = reinterpret_cast<float4*>(rowp + ((c ^ (r & 7)) << 4));
                        const float4 v = *p;
                        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                        if (a.mask_hi) *p = h;
                        hi[4 * c] = v.x; hi[4 * c + 1] = v.y; hi[4 * c + 2] = v.z; hi[4 * c + 3] = v.w;
                        lo[4 * c] = v.x - h.x; lo[4 * c + 1] = v.y - h.y; lo[4 * c + 2] = v.z - h.z; lo[4 * c + 3] = v.w - h.w;
                    }
                } else {                                 // element (k, mn = r): box r / 32, row k, 32-byte chunk ((r % 32) / 8) ^ (k & 3)
                    unsigned char* atom = sa + (r >> 5) * 4096 + (r & 7) * 4;
                    const int ch = (r & 31) >> 3;
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        float* p = reinterpret_cast<float*>(atom + k * 128 + ((ch ^ (k & 3)) << 5));
                        const float v = *p;
                        const float h = tf32_hi(v);
                        if (a.mask_hi) *p = h;
                        hi[k] = v;
                        lo[k] = v - h;
                    }
                }
                tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + alo_col0 + s * a_cols, lo);
                if (a.ts_all) tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + alo_col0 + s * a_cols + 32, hi);
                float4* bh = reinterpret_cast<float4*>(sa + A_BYTES);
                float4* bl = reinterpret_cast<float4*>(sa + A_BYTES + b_bytes);
                const int nchunk = a.b_lo_tma ? 0 : (b_bytes >> 4);
                for (int i = tid; i < nchunk; i += 128) {
                    const float4 v = bh[i];
                    const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                    if (a.mask_hi) bh[i] = h;
                    bl[i] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                }
                fence_proxy_async();                     // generic-proxy writes -> async proxy (UMMA reads)
                tc_fence_before();                       // tcgen05.st (already waited) ordered before the arrive
                __syncwarp();
                if (lane == 0) mbar_arrive(full_split + s);
                if (++s == (uint32_t)a.nstage) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ promoters: TMEM chunks -> fp32 masters -> global
        const int q = warp & 3;
        const int r = q * 32 + lane;
        long long it = 0;
        uint32_t n_stores = 0;                                            // TMA stores issued by this warp (buffer = parity)
        const bool vec_ok = (a.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(a.D) & 15u) == 0 && (a.split_stride & 3) == 0;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            int mt, nt, sp;
            unit_of(u, mt, nt, sp);
            long long k0;
            const int nkb = kblocks(sp, k0);
            const int n_chunks = (nkb + CH - 1) / CH;
            float macc[TN_MAX];
#pragma unroll
            for (int i = 0; i < TN_MAX; ++i) macc[i] = 0.f;
            for (int ch = 0; ch < n_chunks; ++ch, ++it) {
                const uint32_t acc = (uint32_t)(it & 1), acc_ph = (uint32_t)((it >> 1) & 1);
                mbar_wait(acc_full + acc, acc_ph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)a.tn;
#pragma unroll
                for (int c = 0; c < TN_MAX / 32; ++c) {
                    if (c * 32 < a.tn) {
                        float v[32];
                        tmem_ld32(taddr + c * 32, v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) macc[c * 32 + i] += v[i];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + acc);
            }
            const long long row = (long long)mt * TM + r;
            if (a.tma_store) {
                // each warp owns rows [q*32, q*32+32) of the tile: per 32-column chunk, lane = row writes its 128 bytes into
                // a 128-byte-swizzled 4 KB box (16-byte chunk index ^ (row & 7): conflict-free), one lane stores the box by TMA
                unsigned char* wbuf = epi + q * 8192;
#pragma unroll
                for (int c = 0; c < TN_MAX / 32; ++c) {
                    const long long col0 = (long long)nt * a.tn + c * 32;
                    if (c * 32 < a.tn && col0 < a.N && (long long)mt * TM + q * 32 < a.M) {      // warp-uniform
                        if (a.bias) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) if (col0 + i < a.N) macc[c * 32 + i] += __ldg(a.bias + col0 + i);
                        }
                        unsigned char* buf = wbuf + (n_stores++ & 1u) * 4096;
                        if (lane == 0) tma_store_wait_read<1>();       // the store that last read this buffer (two groups ago)
                        __syncwarp();
                        if (a.transpose_out) {
                            // box rows = n (32 columns of the tile), box columns = m (this warp's 32 rows): lane m writes
                            // element (n = i, m) -- 32 consecutive floats per instruction, chunk (m / 4) ^ (n & 7)
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                *reinterpret_cast<float*>(buf + i * 128 + ((((lane >> 2) ^ (i & 7)) << 4) | ((lane & 3) << 2))) = macc[c * 32 + i];
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                *reinterpret_cast<float4*>(buf + lane * 128 + ((i ^ (lane & 7)) << 4)) =
                                    make_float4(macc[c * 32 + 4 * i], macc[c * 32 + 4 * i + 1], macc[c * 32 + 4 * i + 2], macc[c * 32 + 4 * i + 3]);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            if (a.transpose_out) tma_store_3d(&mapD, buf, mt * TM + q * 32, (int)col0, sp);
                            else tma_store_3d(&mapD, buf, (int)col0, mt * TM + q * 32, sp);
                        }
                    }
                }
                continue;
            }
            float* out = a.D + (size_t)sp * a.split_stride + (size_t)row * a.ldd;
#pragma unroll
            for (int c = 0; c < TN_MAX / 32; ++c) {
                const long long col0 = (long long)nt * a.tn + c * 32;
                if (c * 32 < a.tn && row < a.M && col0 < a.N) {
                    if (a.bias) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (col0 + i < a.N) macc[c * 32 + i] += __ldg(a.bias + col0 + i);
                    }
                    if (vec_ok && col0 + 32 <= a.N) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<float4*>(out + col0 + 4 * i) =
                                make_float4(macc[c * 32 + 4 * i], macc[c * 32 + 4 * i + 1], macc[c * 32 + 4 * i + 2], macc[c * 32 + 4 * i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) if (col0 + i < a.N) out[col0 + i] = macc[c * 32 + i];
                    }
                }
            }
        }
    }
    if (a.tma_store && warp >= 6 && lane == 0) tma_store_wait_all();      // shared memory must outlive the bulk stores
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// lo[i] = x[i] - tf32_hi(x[i])   (weights: split once per call instead of once per M tile)
__global__ void __launch_bounds__(256)
split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(x + i);
        lo[i] = v - tf32_hi(v);
    }
}

// out[i] = sum_s partial[s][i] in split order (+ bias): the deterministic end of a split-K product
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int n_split, long long stride, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < n_split; ++s) acc += __ldcs(partial + (size_t)s * stride + i);
        out[i] = acc;
    }
}

struct Plan {
    int tn, n_mt, n_nt, n_split, nstage, tma_store, swap;
    long long k_per_split;
    size_t blo_off;
    size_t smem, ws_bytes;
    uint32_t tmem_cols;
};

static int ts_all_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DFM_G3_TS_ALL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

// mode 2 (weight gradient dW[M][N] = sum_k dY[k][M] X[k][N]) runs SWAPPED when the TMA-store epilogue is available: the
// kernel computes dW^T = X^T dY (rows = N of the caller, columns = M) and stores every tile transposed.  The operand that
// goes through the splitter warps and tensor memory is then X, and dY -- the small one (K x M floats) -- becomes the B
// operand whose lo part is precomputed by one split_lo pass and arrives by TMA, exactly like the weights of modes 0 / 1:
// the in-kernel split of a 16 KB B tile per k-block was what bounded the un-swapped product (tensor pipe 32 %).
static int make_plan(int mode, long long M, long long N, long long K, Plan& p) {
    DFM_REQUIRE(mode >= 0 && mode <= 2 && M > 0 && N > 0 && K > 0, DFM_ERR_INVALID, "dfm_gemm3: bad mode / shape");
    p.swap = 0;
    if (mode == 2 && N % 4 == 0 && M % 4 == 0) {
        const char* e = getenv("DFM_G3_TMA_STORE");
        const char* w = getenv("DFM_G3_DW_SWAP");
        if (!(e && e[0] == '0') && !(w && w[0] == '0')) { p.swap = 1; const long long t = M; M = N; N = t; }
    }
    DFM_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), DFM_ERR_UNSUPPORTED, "dfm_gemm3: dimension too large");
    // TMA needs 16-byte global strides: the contiguous extent of each operand must be a multiple of 4 floats
    const long long a_inner = mode == 2 ? M : K, b_inner = mode == 0 ? K : N;
    DFM_REQUIRE(a_inner % 4 == 0 && b_inner % 4 == 0, DFM_ERR_UNSUPPORTED,
                "dfm_gemm3: contiguous extents (%lld, %lld) must be multiples of 4", a_inner, b_inner);
    p.n_nt = (int)ceil_div(N, TN_MAX);
    p.tn = (int)(ceil_div(ceil_div(N, p.n_nt), 32) * 32);
    p.n_mt = (int)ceil_div(M, TM);
    p.n_split = 1;
    if (mode == 2) {
        // split K so that the persistent grid runs whole waves: units = tiles * splits; a unit count just above a
        // multiple of the SM count leaves most SMs idle for a full unit (40 tiles x 8 splits = 2.16 waves -> 3 rounds)
        const long long base = (long long)p.n_mt * p.n_nt, sms = sm_count();
        long long want = ceil_div(2LL * sms, base);
        const long long max_split = ceil_div(K, 8LL * KB);            // at least 8 k-blocks per split
        if (want > max_split) want = max_split;
        if (want < 1) want = 1;
        long long best = want;
        double best_eff = 0.0;
        for (long long s = (want + 1) / 2; s <= max_split && s <= 3 * want; ++s) {
            const long long units = base * s;
            const double eff = (double)units / (double)(ceil_div(units, sms) * sms);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
            if (eff >= 0.97) break;                                    // the smallest split count that fills its waves
        }
        p.n_split = (int)best;
    }
    p.k_per_split = ceil_div(ceil_div(K, p.n_split), KB) * KB;
    p.n_split = (int)ceil_div(K, p.k_per_split);
    const int stage_bytes = A_BYTES + 2 * p.tn * 128;
    const size_t fixed0 = (3 * MAXST + 4) * 8 + 16 + 1024;
    p.nstage = MAXST;
    const int a_cols = ts_all_enabled() ? 64 : 32;
    {   // TMA-store epilogue: 16-byte global strides (N % 4 == 0); DFM_G3_TMA_STORE=0 keeps the per-thread stores
        const char* e = getenv("DFM_G3_TMA_STORE");
        p.tma_store = ((p.swap ? M : N) % 4 == 0 && !(e && e[0] == '0')) ? 1 : 0;
    }
    const size_t fixed = fixed0 + (p.tma_store ? EPI_BYTES : 0);
    while (p.nstage > 2 && ((size_t)p.nstage * stage_bytes + fixed > 227 * 1024 || 2 * p.tn + p.nstage * a_cols > 512)) --p.nstage;
    p.smem = (size_t)p.nstage * stage_bytes + fixed;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.tn + p.nstage * a_cols)) cols <<= 1;
    p.tmem_cols = cols;
    p.ws_bytes = p.n_split > 1 ? align_up((size_t)p.n_split * M * N * 4, 256) : 0;
    p.blo_off = p.ws_bytes;
    if (mode != 2 || p.swap) p.ws_bytes += align_up((size_t)N * K * 4, 256);      // B_lo of the weight / dY operand
    return DFM_OK;
}

}  // namespace g3
}  // namespace dfm

using namespace dfm;

extern "C" {

size_t dfm_gemm3_workspace_bytes(int mode, int64_t M, int64_t N, int64_t K) {
    g3::Plan p;
    if (g3::make_plan(mode, M, N, K, p) != DFM_OK) return 0;
    return p.ws_bytes + 256;
}

int dfm_gemm3(int mode, const float* A, const float* B, float* D, const float* bias, int64_t M, int64_t N, int64_t K,
              void* workspace, size_t workspace_bytes, void* stream) {
    using namespace g3;
    DFM_REQUIRE(A && B && D, DFM_ERR_INVALID, "dfm_gemm3: null operand");
    Plan p;
    int rc = make_plan(mode, M, N, K, p);
    if (rc) return rc;
    DFM_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15u) == 0, DFM_ERR_UNSUPPORTED,
                "dfm_gemm3: operands must be 16-byte aligned");
    DFM_REQUIRE(p.ws_bytes == 0 || (workspace && workspace_bytes >= p.ws_bytes), DFM_ERR_WORKSPACE,
                "dfm_gemm3: workspace %zu < %zu", workspace_bytes, p.ws_bytes);
    DFM_REQUIRE(p.n_split == 1 || !bias, DFM_ERR_UNSUPPORTED, "dfm_gemm3: bias with split-K is not supported");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p.swap) {                       // internal product: rows = the caller's N (operand B), columns = the caller's M (operand A)
        const float* t = A; A = B; B = t;
        const int64_t m = M; M = N; N = m;
    }
    Args a;
    memset(&a, 0, sizeof(a));
    a.transpose_out = p.swap;
    a.M = M; a.N = N; a.K = K; a.tn = p.tn; a.n_mt = p.n_mt; a.n_nt = p.n_nt; a.n_split = p.n_split;
    a.k_per_split = p.k_per_split; a.nstage = p.nstage; a.tmem_cols = p.tmem_cols;
    a.a_mn = mode == 2 ? 1 : 0; a.b_mn = mode == 0 ? 0 : 1;
    a.mask_hi = getenv("DFM_G3_MASK") ? 1 : 0;
    a.b_lo_tma = ((mode != 2 || p.swap) && !getenv("DFM_G3_SPLIT_B_IN_KERNEL")) ? 1 : 0;
    a.ts_all = ts_all_enabled();
    a.tma_store = p.tma_store;
    a.bias = bias;
    // (transposed output: the stored matrix is [N][M], row stride M)
    if (p.n_split > 1) { a.D = static_cast<float*>(workspace); a.ldd = p.swap ? M : N; a.split_stride = M * N; }
    else { a.D = D; a.ldd = p.swap ? M : N; a.split_stride = 0; }
    CUtensorMap mapA, mapB;
    // K-major operand [rows][K]: box {32 k, rows};  MN-major operand [K][MN]: box {32 mn, 32 k}
    rc = a.a_mn ? tc::make_tmap_2d(&mapA, A, K, M, 32, true) : tc::make_tmap_2d(&mapA, A, M, K, TM);
    if (rc) return rc;
    rc = a.b_mn ? tc::make_tmap_2d(&mapB, B, K, N, 32, true) : tc::make_tmap_2d(&mapB, B, N, K, p.tn);
    if (rc) return rc;
    CUtensorMap mapBlo = mapB;
    if (a.b_lo_tma) {
        float* blo = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.blo_off);   // behind the split-K partials
        const long long nb = N * K;
        long long blocks = ceil_div(nb, 256 * 4);
        if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
        split_lo_kernel<<<(unsigned)blocks, 256, 0, st>>>(B, blo, nb);
        rc = a.b_mn ? tc::make_tmap_2d(&mapBlo, blo, K, N, 32, true) : tc::make_tmap_2d(&mapBlo, blo, N, K, p.tn);
        if (rc) return rc;
    }
    CUtensorMap mapD = mapA;
    if (a.tma_store) {
        DFM_REQUIRE((reinterpret_cast<uintptr_t>(a.D) & 15u) == 0, DFM_ERR_UNSUPPORTED, "dfm_gemm3: output must be 16-byte aligned");
        rc = p.swap ? tc::make_tmap_out_3d(&mapD, a.D, p.n_split, N, M, a.ldd, a.split_stride)
                    : tc::make_tmap_out_3d(&mapD, a.D, p.n_split, M, N, a.ldd, a.split_stride);
        if (rc) return rc;
    }
    DFM_CHECK_CUDA(cudaFuncSetAttribute(gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    long long grid = (long long)p.n_mt * p.n_nt * p.n_split;
    if (grid > sm_count()) grid = sm_count();
    gemm3_kernel<<<(unsigned)grid, THREADS, p.smem, st>>>(a, mapA, mapB, mapBlo, mapD);
    if (p.n_split > 1) {
        const long long n = M * N;
        long long blocks = ceil_div(n, 256);
        if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
        splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(workspace), p.n_split, M * N, D, n);
    }
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
