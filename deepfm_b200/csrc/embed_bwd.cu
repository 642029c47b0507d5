// K2: backward of K1 for every FeatureEmbedding parameter, fused with the FM backward and with
// the L2 penalty gradient (reference: autograd through embedding.py:76-126 + fm.py:18-23, and
// base.py:78-83; ATen embedding_dense_backward / _embedding_bag_dense_backward / mm).
//
//   sort      stable LSB radix sort of (key = global row, payload = b*S + slot), PAD keys last
//   segreduce one lane group per chunk of sorted positions; the upstream gradient of a slot
//             g_raw = g_flat + P^T (g_field + g_fm (fm_sum - e)) is rebuilt on the fly, so the
//             FM gradient never exists in HBM; segments inside a chunk are written directly
//             (+ 2*l2*w[row]); the open head/tail partial sums go to a side buffer
//   stitch    segments that span units (their owner units are listed by segreduce): one lane group each
//             adds the units' partials in unit order; hot rows that span many units go to a second
//             kernel, one block each, with a fixed strided order
//   pgrads    DENSE-field Linear and projection gradients: per-slice partial sums in shared
//             memory, then a fixed-order reduction over slices (+ 2*l2*p)
// No float atomics anywhere: every output element has exactly one writer and a fixed summation
// order, so the gradients are bit-reproducible run to run.
#include <stdlib.h>

#include <type_traits>

#include <cub/device/device_radix_sort.cuh>

#include "plan.cuh"

namespace dfm {

constexpr int CHUNK = 32;      // sorted positions per lane group in segreduce
constexpr int SEG2_UNIT = 128; // sorted positions per span of seg2_kernel (the carry records are sized for it)
constexpr int PG_TILE = 32;    // samples per shared-memory tile in pgrads
constexpr int PG_THREADS = 256;

static inline int slot_bits_of(int S) { int b = 0; while ((1 << b) < S) ++b; return b; }

struct BwdArgs {
    const float* g_first;
    const float* g_field;
    const float* g_flat;
    const float* g_fm;
    const float* fe;      // field embeddings (B, F, D)
    const float* flat;    // raw concat (B, T)
    const float* fm_sum;  // (B, D)
    const uint32_t* aux;
    const float* l2_gscale;
    float l2x2;           // 2 * lambda
    int mode;
    long long B, N;       // N = B * S
    const uint32_t* skeys;
    const uint32_t* spay;
    float* row_grad2;
    float* row_grad1;
    float* head2; float* head1; float* tail2; float* tail1;  // per-unit open partial sums
    long long* tail_start;
    float* headg; float* tailg;          // per-unit open partial sums of g_fm (plain SPARSE fields)
    int* tail_field;                      // field of the segment that leaves the unit
    int direct;                           // 1: payload = row index into g_flat (M, row_stride); field from key
    int row_stride;                       // direct mode: floats per row = max_tdim + 4 ([g, g_first, g_fm, 0, 0])
    int slot_bits;                        // payload = (b << slot_bits) | slot
    unsigned long long* counters;  // {n_valid, n_unique}
    unsigned* open_count;          // number of units that own a segment leaving the unit
    unsigned* open_list;           // those units (order irrelevant: one writer per segment)
    unsigned* long_count;          // segments spanning more than LONG_SPAN units: (owner unit, last unit) pairs
    unsigned* long_list;
    unsigned* span_counter;        // seg2: next unclaimed span (dynamic scheduling)
    const float* g_sc;             // direct mode, split layout: g_flat = (M, tdim) vectors, g_sc = (M, 4) [g_first, g_fm, 0, 0]
    unsigned pad_key;              // PAD key of the sorted stream (normally the plan's; owner-major keys have their own)
    // peer output (row-sharded tables, sample side): the finished segment sums are not table gradients but rows of the
    // gradient EXCHANGE -- segment u (its ordinal among the segment heads, uidx[head position]) belongs to the owner
    // whose range [peer_start[o], peer_start[o + 1]) contains u and is stored straight into that GPU's buffers
    // (NVLink peer stores): vector at peer_vec[o] + (u - start) * tdim, scalars [sum g_first, sum g_fm, 0, 0] at
    // peer_sc[o] + (u - start) * 4, everything times peer_scale.  No table row is read, no L2 term added (the owner does).
    int peer_n;
    long long peer_start[17];
    float* peer_vec[16];
    float* peer_sc[16];
    const uint32_t* uidx;
    float peer_scale;
};

// destination of exchange row u (peer output mode)
__device__ __forceinline__ void peer_dst(const BwdArgs& a, long long u, int tdim, float*& vec, float*& sc) {
    int o = 0;
#pragma unroll 1
    for (int q = 1; q < a.peer_n; ++q) if (u >= a.peer_start[q]) o = q;
    const long long r = u - a.peer_start[o];
    vec = a.peer_vec[o] + (size_t)r * tdim;
    sc = a.peer_sc[o] + (size_t)r * 4;
}

// payload of key position i = b*S + slot:  (b << bits) | slot   (decoded with a shift and a mask)
__global__ void payload_kernel(uint32_t* p, long long n, int S, int bits) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(i / S);
        p[i] = (b << bits) | (uint32_t)(i - (long long)b * S);
    }
}

// g[i] = coef * p[i]   (dense-mode L2 gradient for every row; coef may be 0 -> zero fill)
__global__ void l2_fill_kernel(const float* __restrict__ p, float* __restrict__ g, long long n,
                               float l2x2, const float* __restrict__ gscale) {
    const float coef = l2x2 * (gscale ? __ldg(gscale) : 1.f);
    const long long n4 = n >> 2;
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (al) {
        if (coef == 0.f) {
            for (; i < n4; i += stride) __stcs(reinterpret_cast<float4*>(g) + i, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
            for (; i < n4; i += stride) {
                float4 w = __ldcs(reinterpret_cast<const float4*>(p) + i);
                __stcs(reinterpret_cast<float4*>(g) + i, make_float4(coef * w.x, coef * w.y, coef * w.z, coef * w.w));
            }
        }
        for (long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
            g[t] = coef == 0.f ? 0.f : coef * p[t];
    } else {
        for (; i < n; i += stride) g[i] = coef == 0.f ? 0.f : coef * p[i];
    }
}

// The same fill for a LIST of small tensors in one launch (blockIdx.y = tensor): dense-mode plans with many small
// tables (the replicated tables of the sharded path: 28 tensors) paid one launch per tensor.
struct L2FillList { const float* p[64]; float* g[64]; long long n[64]; };

__global__ void l2_fill_multi_kernel(const __grid_constant__ L2FillList L, float l2x2, const float* __restrict__ gscale) {
    const float coef = l2x2 * (gscale ? __ldg(gscale) : 1.f);
    const float* __restrict__ p = L.p[blockIdx.y];
    float* __restrict__ g = L.g[blockIdx.y];
    const long long n = L.n[blockIdx.y];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15u) == 0;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (al) {
        const long long n4 = n >> 2;
        for (; i < n4; i += stride) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (coef != 0.f) w = __ldcs(reinterpret_cast<const float4*>(p) + i);
            __stcs(reinterpret_cast<float4*>(g) + i, make_float4(coef * w.x, coef * w.y, coef * w.z, coef * w.w));
        }
        for (i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g[i] = coef != 0.f ? coef * p[i] : 0.f;
    } else {
        for (; i < n; i += stride) g[i] = coef != 0.f ? coef * p[i] : 0.f;
    }
}

// ---- upstream gradient of one id slot, chunk c (V floats) of the raw embedding row
template <int V>
__device__ __forceinline__ VecF<V> slot_grad(const DevPlan& P, const BwdArgs& a, const FieldDev& fd,
                                             int f, long long b, int l, int c) {
    const int D = P.D;
    VecF<V> g = vzero<V>();
    if (a.g_flat) g = vload_stream<V>(a.g_flat + (size_t)b * P.T + fd.flat_off + c * V);
    const float gfm = a.g_fm ? __ldg(a.g_fm + b) : 0.f;
    const size_t eoff = ((size_t)b * P.n_fields + f) * D;
    if (fd.proj == nullptr) {
        if (a.g_field) {
            VecF<V> t = vload_stream<V>(a.g_field + eoff + c * V);
#pragma unroll
            for (int v = 0; v < V; ++v) g.v[v] += t.v[v];
        }
        if (a.g_fm) {
            VecF<V> s = vload<V>(a.fm_sum + (size_t)b * D + c * V);
            VecF<V> e = vload_stream<V>(a.fe + eoff + c * V);
#pragma unroll
            for (int v = 0; v < V; ++v) g.v[v] = fmaf(gfm, s.v[v] - e.v[v], g.v[v]);
        }
    } else if (a.g_field || a.g_fm) {
        const int d = fd.dim;
        for (int k = 0; k < D; ++k) {
            float gek = a.g_field ? __ldg(a.g_field + eoff + k) : 0.f;
            if (a.g_fm) gek = fmaf(gfm, __ldg(a.fm_sum + (size_t)b * D + k) - __ldg(a.fe + eoff + k), gek);
#pragma unroll
            for (int v = 0; v < V; ++v) g.v[v] = fmaf(gek, __ldg(fd.proj + (size_t)k * d + c * V + v), g.v[v]);
        }
    }
    if (fd.kind == DFM_SEQUENCE) {
        if (fd.combiner == DFM_MEAN) {
            const float inv = __uint_as_float(__ldg(a.aux + (size_t)b * P.A + fd.aux_off));
#pragma unroll
            for (int v = 0; v < V; ++v) g.v[v] *= inv;
        } else if (fd.combiner == DFM_MAX) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                if ((int)__ldg(a.aux + (size_t)b * P.A + fd.aux_off + c * V + v) != l) g.v[v] = 0.f;
        }
    }
    return g;
}

__device__ __forceinline__ float slot_grad1(const DevPlan& P, const BwdArgs& a, const FieldDev& fd,
                                            long long b, int l) {
    float g = a.g_first ? __ldg(a.g_first + b) : 0.f;
    if (fd.kind == DFM_SEQUENCE) {
        if (fd.combiner == DFM_MEAN) g *= __uint_as_float(__ldg(a.aux + (size_t)b * P.A + fd.aux_off));
        else if (fd.combiner == DFM_MAX && (int)__ldg(a.aux + (size_t)b * P.A + fd.aux_off + fd.dim) != l) g = 0.f;
    }
    return g;
}

// Per-field record staged in shared memory by segreduce / stitch (lane-divergent reads of the
// by-value plan would serialise in the constant bank).
struct FieldB {
    const float* w2;
    const float* w1;
    unsigned row_base;
    int dim, flat_off, aux_off;
    int flags;              // kind | combiner << 4 | plain << 8   (plain: SPARSE without projection)
};

__device__ __forceinline__ void stage_fields(const DevPlan& P, FieldB* t) {
    for (int f = threadIdx.x; f < P.n_fields; f += blockDim.x) {
        const FieldDev& fd = P.f[f];
        FieldB e;
        e.w2 = fd.w2; e.w1 = fd.w1; e.row_base = (unsigned)fd.row_base; e.dim = fd.dim;
        e.flat_off = fd.flat_off; e.aux_off = fd.aux_off;
        e.flags = fd.kind | (fd.combiner << 4) | ((fd.kind == DFM_SPARSE && fd.proj == nullptr) ? 0x100 : 0) | (fd.foreign ? 0x200 : 0);
        t[f] = e;
    }
}

// ---- write the finished gradient of one row (segment)
// For plain SPARSE fields the FM term sum_b g_fm[b] (S[b] - e[b,f]) has e[b,f] == w[row] for every
// member of the segment, so the kernel only accumulates sum_b g_fm[b] S[b] and gs = sum_b g_fm[b]
// and folds  -gs * w[row]  in here, together with the L2 term  coef * w[row]:  the field
// embeddings are never re-read.
template <int V>
__device__ __forceinline__ void write_row(const DevPlan& P, const DevGrads& GR, const BwdArgs& a, float coef,
                                          const FieldB& fb, uint32_t key, int f, long long head_pos, int j,
                                          const VecF<V>& acc, float acc1, float gs,
                                          bool have_pre = false, VecF<V> wpre = VecF<V>(), float w1pre = 0.f) {
    if (a.peer_n > 0) {     // exchange row instead of a table gradient (see BwdArgs)
        float *pv, *ps;
        peer_dst(a, (long long)__ldg(a.uidx + head_pos), P.max_tdim, pv, ps);
        if (j < P.max_tdim / V) {
            VecF<V> out = acc;
#pragma unroll
            for (int v = 0; v < V; ++v) out.v[v] *= a.peer_scale;
            vstore<V>(pv + j * V, out);
        }
        if (j == 0) *reinterpret_cast<float4*>(ps) = make_float4(acc1 * a.peer_scale, gs * a.peer_scale, 0.f, 0.f);
        return;
    }
    const long long row = (long long)(key - fb.row_base);
    if (j < fb.dim / V) {   // table dims are <= G * V (checked on the host)
        VecF<V> out = acc;
        const float cw = coef - gs;
        if (cw != 0.f) {
            const VecF<V> w = have_pre ? wpre : vload<V>(fb.w2 + (size_t)row * fb.dim + j * V);
#pragma unroll
            for (int v = 0; v < V; ++v) out.v[v] = fmaf(cw, w.v[v], out.v[v]);
        }
        if (a.mode == DFM_GRAD_DENSE) vstore<V>(GR.g[f].gw2 + (size_t)row * fb.dim + j * V, out);
        else vstore<V>(a.row_grad2 + (size_t)head_pos * P.max_tdim + j * V, out);
    }
    if (j == 0) {
        float o1 = acc1;
        if (coef != 0.f) o1 = fmaf(coef, have_pre ? w1pre : __ldg(fb.w1 + row), o1);
        if (a.mode == DFM_GRAD_DENSE) GR.g[f].gw1[row] = o1;
        else a.row_grad1[head_pos] = o1;
    }
}

template <int V>
struct ChunkState {
    int flag, f, n_heads, n_valid;
    uint32_t cur;
    long long seg_start;
    VecF<V> acc, w;      // w: prefetched table row chunk of the current segment (plain path)
    float acc1, gs, w1;
};

// Shared-memory image of one segreduce block (one "unit" of gpb * CHUNK sorted positions).
struct SegSmem {
    float *sH, *sT, *sH1, *sT1, *sHg, *sTg;     // per-chunk open partial sums
    int* sFlag;
    unsigned long long* s_goff;                 // per position: element offset of its gradient row
    const float** s_wptr;                       // per position: &w2[row][0] and &w1[row] of its table row
    const float** s_w1ptr;
    uint32_t *s_keys, *s_pay, *s_b;             // sorted key (+1 look-ahead), payload, sample index
    float *s_m, *s_o;                           // per position: g_fm[b], first-order upstream gradient
    unsigned short *s_f, *s_slotf, *s_slotl;    // per position field; slot -> field / bag position
    FieldB* t_field;
};

__host__ __device__ inline size_t seg_smem_bytes(int gpb, int tdim, int unit, int S, int F) {
    size_t n = (size_t)gpb * (2 * tdim + 4) * 4 + (size_t)gpb * 4;   // partials + flags
    n = (n + 7) & ~(size_t)7;
    n += (size_t)unit * 8 * 3;                                       // s_goff, s_wptr, s_w1ptr
    n += (size_t)(unit + 1) * 4 + (size_t)unit * 4 * 4;              // keys, pay, b, m, o
    n += (size_t)(((unit + 1) & ~1) + 2 * ((S + 1) & ~1)) * 2;       // s_f, slot tables
    n = (n + 15) & ~(size_t)15;
    return n + (size_t)F * sizeof(FieldB);
}

__device__ __forceinline__ void seg_carve(unsigned char* base, int gpb, int tdim, int unit, int S, SegSmem& m) {
    float* fp = reinterpret_cast<float*>(base);
    m.sH = fp; m.sT = m.sH + gpb * tdim; m.sH1 = m.sT + gpb * tdim; m.sT1 = m.sH1 + gpb;
    m.sHg = m.sT1 + gpb; m.sTg = m.sHg + gpb;
    m.sFlag = reinterpret_cast<int*>(m.sTg + gpb);
    uintptr_t q = (reinterpret_cast<uintptr_t>(m.sFlag + gpb) + 7) & ~(uintptr_t)7;
    m.s_goff = reinterpret_cast<unsigned long long*>(q);
    m.s_wptr = reinterpret_cast<const float**>(m.s_goff + unit);
    m.s_w1ptr = m.s_wptr + unit;
    m.s_keys = reinterpret_cast<uint32_t*>(m.s_w1ptr + unit);
    m.s_pay = m.s_keys + unit + 1;
    m.s_b = m.s_pay + unit;
    m.s_m = reinterpret_cast<float*>(m.s_b + unit);
    m.s_o = m.s_m + unit;
    m.s_f = reinterpret_cast<unsigned short*>(m.s_o + unit);
    m.s_slotf = m.s_f + ((unit + 1) & ~1);
    m.s_slotl = m.s_slotf + ((S + 1) & ~1);
    m.t_field = reinterpret_cast<FieldB*>((reinterpret_cast<uintptr_t>(m.s_slotl + ((S + 1) & ~1)) + 15) & ~(uintptr_t)15);
}

struct SegCtx {
    int gl, j, tdim, nlane, c0;
    long long p0, unit0;
    float coef;
    int bits;
    uint32_t smask;
};

// One chunk of CHUNK sorted positions.  Per-position metadata (row offset, sample, g_fm, first-order
// gradient, field) was decoded ONCE per block into shared memory, so an item costs two address
// computes and two 128-bit loads per lane.  NB items are loaded before any is consumed (all loads of
// a batch in flight together), then the segment logic runs over them in sorted order.
// GENERIC = true (NB = 1) additionally handles sequence-bag / projected fields.
template <int V, int NB, bool GENERIC, bool HAS_FIELD>
__device__ __forceinline__ void process_chunk(const DevPlan& P, const DevGrads& GR, const BwdArgs& a,
                                              const SegSmem& m, const SegCtx& cx, ChunkState<V>& st) {
    const int j = cx.j, gl = cx.gl, tdim = cx.tdim, nlane = cx.nlane, c0 = cx.c0;
    const uint32_t PAD = P.pad_key;
    const bool has_fm = a.g_fm != nullptr;
    const int cend = (int)((a.N - cx.p0 < CHUNK) ? a.N - cx.p0 : CHUNK);
    uint32_t prev = PAD;
    if (c0 > 0) prev = m.s_keys[c0 - 1];
    else if (cx.unit0 > 0) prev = __ldg(a.skeys + cx.unit0 - 1);
    bool started_before = prev == st.cur;
    st.n_heads = started_before ? 0 : 1;
    bool ended = false;
    const bool need_w = has_fm || cx.coef != 0.f || a.direct;
    for (int pb = 0; pb < cend && !ended; pb += NB) {
        uint32_t k4[NB];
        VecF<V> gA[NB], gB[HAS_FIELD ? NB : 1], sv4[NB], wv4[GENERIC ? 1 : NB];
        float m4[NB], o4[NB], w14[GENERIC ? 1 : NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int idx = c0 + pb + i;
            k4[i] = (pb + i < cend) ? m.s_keys[idx] : PAD;
            gA[i] = vzero<V>(); sv4[i] = vzero<V>(); m4[i] = 0.f; o4[i] = 0.f;
            if (HAS_FIELD) gB[i] = vzero<V>();
            if (!GENERIC) { wv4[i] = vzero<V>(); w14[i] = 0.f; }
            if (k4[i] == PAD) continue;
            bool plain = true;
            int dimf = tdim;
            if (GENERIC) {
                const FieldB& fb = m.t_field[m.s_f[idx]];
                plain = (fb.flags & 0x100) != 0;
                dimf = fb.dim;
            }
            if (plain) {
                const bool on = j < dimf / V;
                if (on && a.g_flat) gA[i] = vload_stream<V>(a.g_flat + m.s_goff[idx] + j * V);
                if (HAS_FIELD && on && a.g_field)
                    gB[i] = vload_stream<V>(a.g_field + ((size_t)m.s_b[idx] * P.n_fields + m.s_f[idx]) * P.D + j * V);
                if (!GENERIC) {   // table row of this key, for the -(sum g_fm) w and 2 l2 w terms at the segment end
                    if (on && need_w) wv4[i] = vload<V>(m.s_wptr[idx] + j * V);
                    if (j == 0 && cx.coef != 0.f) w14[i] = __ldg(m.s_w1ptr[idx]);
                }
                if (has_fm) {
                    m4[i] = m.s_m[idx];
                    if (on) sv4[i] = vload<V>(a.fm_sum + (size_t)m.s_b[idx] * P.D + j * V);
                }
                o4[i] = m.s_o[idx];
            } else {
                const int ff = m.s_f[idx];
                const FieldDev& fd = P.f[ff];
                const uint32_t pay = m.s_pay[idx];
                const int l = m.s_slotl[pay & cx.smask];
                const long long b = pay >> cx.bits;
                if (j < fd.dim / V) gA[i] = slot_grad<V>(P, a, fd, ff, b, l, j);
                o4[i] = slot_grad1(P, a, fd, b, l);
            }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (k4[i] == PAD) { ended = true; break; }
            if (k4[i] != st.cur) {   // previous segment ended inside this chunk
                if (started_before) {
                    if (j < nlane) vstore<V>(m.sH + gl * tdim + j * V, st.acc);
                    if (j == 0) { m.sH1[gl] = st.acc1; m.sHg[gl] = st.gs; }
                    st.flag |= 1;
                } else {
                    write_row<V>(P, GR, a, cx.coef, m.t_field[st.f], st.cur, st.f, st.seg_start, j, st.acc, st.acc1, st.gs, !GENERIC, st.w, st.w1);
                }
                st.cur = k4[i]; st.seg_start = cx.p0 + pb + i; started_before = false; ++st.n_heads;
                st.f = m.s_f[c0 + pb + i];
                st.acc = vzero<V>(); st.acc1 = 0.f; st.gs = 0.f;
            }
            if (!GENERIC) { st.w = wv4[i]; st.w1 = w14[i]; }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                float g = gA[i].v[v];
                if (HAS_FIELD) g += gB[i].v[v];
                st.acc.v[v] += fmaf(m4[i], sv4[i].v[v], g);
            }
            st.gs += m4[i];
            st.acc1 += o4[i];
            ++st.n_valid;
        }
    }
    const bool continues = !ended && cend == CHUNK && m.s_keys[c0 + CHUNK] == st.cur;
    if (started_before) {
        if (j < nlane) vstore<V>(m.sH + gl * tdim + j * V, st.acc);
        if (j == 0) { m.sH1[gl] = st.acc1; m.sHg[gl] = st.gs; }
        st.flag |= 1 | (continues ? 2 : 0);
    } else if (continues) {
        if (j < nlane) vstore<V>(m.sT + gl * tdim + j * V, st.acc);
        if (j == 0) { m.sT1[gl] = st.acc1; m.sTg[gl] = st.gs; }
        st.flag |= 4;
    } else {
        write_row<V>(P, GR, a, cx.coef, m.t_field[st.f], st.cur, st.f, st.seg_start, j, st.acc, st.acc1, st.gs, !GENERIC, st.w, st.w1);
    }
}

// Fast path of process_chunk for units whose fields are all plain SPARSE, with g_flat present and
// every lane owning a chunk (max_tdim == G * V).  Every load is UNCONDITIONAL (invalid items re-read
// the chunk's first position and are never consumed) and nothing derived from a loaded value is
// touched before all loads of the batch are issued; the table row of the current segment is taken
// from the batch registers by compile-time index, so ptxas keeps NB x (2..4) 128-bit loads in flight.
template <int V, int NB, bool HAS_FIELD>
__device__ __forceinline__ void process_chunk_fast(const DevPlan& P, const DevGrads& GR, const BwdArgs& a,
                                                   const SegSmem& m, const SegCtx& cx, ChunkState<V>& st) {
    const int j = cx.j, gl = cx.gl, tdim = cx.tdim, c0 = cx.c0;
    const uint32_t PAD = P.pad_key;
    const bool has_fm = a.g_fm != nullptr;
    const bool need_w = has_fm || cx.coef != 0.f || a.direct;
    const bool need_w1 = cx.coef != 0.f;
    const int cend = (int)((a.N - cx.p0 < CHUNK) ? a.N - cx.p0 : CHUNK);
    uint32_t prev = PAD;
    if (c0 > 0) prev = m.s_keys[c0 - 1];
    else if (cx.unit0 > 0) prev = __ldg(a.skeys + cx.unit0 - 1);
    bool started_before = prev == st.cur;
    st.n_heads = started_before ? 0 : 1;
    bool ended = false;
    for (int pb = 0; pb < cend && !ended; pb += NB) {
        uint32_t k4[NB];
        VecF<V> gA[NB], gB[NB], sv4[NB], wv4[NB];
        float m4[NB], o4[NB], w14[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int real = c0 + pb + i;
            k4[i] = (pb + i < cend) ? m.s_keys[real] : PAD;
            const int idx = k4[i] == PAD ? c0 : real;      // always a valid, decoded position
            gA[i] = vload_stream<V>(a.g_flat + m.s_goff[idx] + j * V);
            if (HAS_FIELD) gB[i] = vload_stream<V>(a.g_field + ((size_t)m.s_b[idx] * P.n_fields + m.s_f[idx]) * P.D + j * V);
            if (has_fm) sv4[i] = vload<V>(a.fm_sum + (size_t)m.s_b[idx] * P.D + j * V);
            if (need_w) wv4[i] = vload<V>(m.s_wptr[idx] + j * V);
            if (need_w1) w14[i] = __ldg(m.s_w1ptr[idx]);
            m4[i] = m.s_m[idx];
            o4[i] = m.s_o[idx];
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            if (k4[i] == PAD) {
                if (i > 0) { if (need_w) st.w = wv4[i > 0 ? i - 1 : 0]; if (need_w1) st.w1 = w14[i > 0 ? i - 1 : 0]; }
                ended = true;
                break;
            }
            if (k4[i] != st.cur) {   // previous segment ended inside this chunk
                if (started_before) {
                    vstore<V>(m.sH + gl * tdim + j * V, st.acc);
                    if (j == 0) { m.sH1[gl] = st.acc1; m.sHg[gl] = st.gs; }
                    st.flag |= 1;
                } else {
                    VecF<V> wrow = st.w;
                    float w1row = st.w1;
                    if (i > 0) { wrow = wv4[i > 0 ? i - 1 : 0]; w1row = w14[i > 0 ? i - 1 : 0]; }
                    write_row<V>(P, GR, a, cx.coef, m.t_field[st.f], st.cur, st.f, st.seg_start, j, st.acc, st.acc1, st.gs,
                                 true, wrow, w1row);
                }
                st.cur = k4[i]; st.seg_start = cx.p0 + pb + i; started_before = false; ++st.n_heads;
                st.f = m.s_f[c0 + pb + i];
                st.acc = vzero<V>(); st.acc1 = 0.f; st.gs = 0.f;
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                float g = gA[i].v[v];
                if (HAS_FIELD) g += gB[i].v[v];
                if (has_fm) g = fmaf(m4[i], sv4[i].v[v], g);
                st.acc.v[v] += g;
            }
            st.gs += m4[i];
            st.acc1 += o4[i];
            ++st.n_valid;
        }
        if (!ended) { if (need_w) st.w = wv4[NB - 1]; if (need_w1) st.w1 = w14[NB - 1]; }
    }
    const bool continues = !ended && cend == CHUNK && m.s_keys[c0 + CHUNK] == st.cur;
    if (started_before) {
        vstore<V>(m.sH + gl * tdim + j * V, st.acc);
        if (j == 0) { m.sH1[gl] = st.acc1; m.sHg[gl] = st.gs; }
        st.flag |= 1 | (continues ? 2 : 0);
    } else if (continues) {
        vstore<V>(m.sT + gl * tdim + j * V, st.acc);
        if (j == 0) { m.sT1[gl] = st.acc1; m.sTg[gl] = st.gs; }
        st.flag |= 4;
    } else {
        write_row<V>(P, GR, a, cx.coef, m.t_field[st.f], st.cur, st.f, st.seg_start, j, st.acc, st.acc1, st.gs, true, st.w, st.w1);
    }
}

// field of a global row key (direct mode: the payload carries no slot)
__device__ __forceinline__ int field_of_key(const FieldB* t, int n_fields, uint32_t key) {
    int f = 0;
    for (int i = 0; i < n_fields; ++i)
        if (t[i].dim > 0 && (t[i].flags & 0xf) != DFM_DENSE && !(t[i].flags & 0x200) && key >= t[i].row_base) f = i;
    return f;
}

// Shared-memory image of one chunk's open partial sums:
//   flag bit0: the chunk's first segment started in an earlier chunk; its partial sum is in sH
//        bit1: that segment covers the whole chunk AND continues into the next one ("through")
//        bit2: the chunk's last segment starts here and continues into the next chunk; sum in sT
template <int V, bool ANY_GENERIC>
__global__ void __launch_bounds__(256)
segreduce_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ DevGrads GR,
                 const __grid_constant__ BwdArgs a, int G) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G;
    const int j = threadIdx.x - gl * G;
    const int tdim = P.max_tdim, nlane = tdim / V;
    const int unit = gpb * CHUNK, S = P.S;
    SegSmem m;
    seg_carve(sm_raw, gpb, tdim, unit, S, m);
    const long long unit0 = (long long)blockIdx.x * unit;   // first position of this block
    const uint32_t PAD = P.pad_key;
    const int bits = a.slot_bits;
    const uint32_t smask = (1u << bits) - 1u;
    for (int i = threadIdx.x; i <= unit; i += blockDim.x) {
        const long long p = unit0 + i;
        m.s_keys[i] = p < a.N ? __ldg(a.skeys + p) : PAD;
        if (i < unit) m.s_pay[i] = p < a.N ? __ldg(a.spay + p) : 0u;
    }
    for (int s = threadIdx.x; s < S; s += blockDim.x) { m.s_slotf[s] = P.slot_field[s]; m.s_slotl[s] = P.slot_pos[s]; }
    stage_fields(P, m.t_field);
    __syncthreads();
    // decode every position once: gradient-row offset, sample, g_fm, first-order gradient, field
    int generic_here = 0;
    for (int i = threadIdx.x; i < unit; i += blockDim.x) {
        const uint32_t key = m.s_keys[i];
        if (key == PAD) continue;
        const uint32_t pay = m.s_pay[i];
        if (a.direct) {
            const int fd_ = field_of_key(m.t_field, P.n_fields, key);
            const FieldB& fk = m.t_field[fd_];
            m.s_f[i] = (unsigned short)fd_;
            m.s_wptr[i] = fk.w2 + (size_t)(key - fk.row_base) * fk.dim;
            m.s_w1ptr[i] = fk.w1 + (key - fk.row_base);
            m.s_goff[i] = (unsigned long long)pay * a.row_stride;
            m.s_b[i] = 0;
            m.s_o[i] = __ldg(a.g_flat + (size_t)pay * a.row_stride + tdim);        // packed first-order gradient
            m.s_m[i] = __ldg(a.g_flat + (size_t)pay * a.row_stride + tdim + 1);    // packed g_fm (for -(sum g_fm) w)
        } else {
            const uint32_t b = pay >> bits;
            const int f = m.s_slotf[pay & smask];
            const FieldB& fb = m.t_field[f];
            m.s_f[i] = (unsigned short)f;
            m.s_wptr[i] = fb.w2 + (size_t)(key - fb.row_base) * fb.dim;
            m.s_w1ptr[i] = fb.w1 + (key - fb.row_base);
            m.s_goff[i] = (unsigned long long)b * P.T + fb.flat_off;
            m.s_b[i] = b;
            m.s_m[i] = a.g_fm ? __ldg(a.g_fm + b) : 0.f;
            m.s_o[i] = a.g_first ? __ldg(a.g_first + b) : 0.f;
            if (!(fb.flags & 0x100)) generic_here = 1;
        }
    }
    // does any position of this unit need the generic gradient (sequence bag / projected field)?
    // Sorted keys group by field, so units are nearly always uniform.
    const int any_generic = ANY_GENERIC ? __syncthreads_or(generic_here) : (__syncthreads(), 0);

    const int c0 = gl * CHUNK;
    const long long p0 = unit0 + c0;
    const float coef = a.l2x2 * (a.l2_gscale ? __ldg(a.l2_gscale) : 1.f);
    float* sH = m.sH; float* sT = m.sT; float* sH1 = m.sH1; float* sT1 = m.sT1; float* sHg = m.sHg; float* sTg = m.sTg;
    int* sFlag = m.sFlag;
    const FieldB* t_field = m.t_field;

    ChunkState<V> st;
    st.flag = 0; st.f = 0; st.n_heads = 0; st.n_valid = 0; st.cur = m.s_keys[c0]; st.seg_start = p0;
    st.acc = vzero<V>(); st.acc1 = 0.f; st.gs = 0.f; st.w = vzero<V>(); st.w1 = 0.f;
    if (st.cur != PAD) {
        st.f = m.s_f[c0];
        const SegCtx cx{gl, j, tdim, nlane, c0, p0, unit0, coef, bits, smask};
        const bool fast = !(ANY_GENERIC && any_generic) && a.g_flat != nullptr && nlane == G;
        if (!fast) process_chunk<V, 1, true, true>(P, GR, a, m, cx, st);
        else if (a.g_field) process_chunk_fast<V, 2, true>(P, GR, a, m, cx, st);
        else process_chunk_fast<V, 4, false>(P, GR, a, m, cx, st);
    }
    const int flag = st.flag, f = st.f, n_heads = st.n_heads, n_valid = st.n_valid;
    const uint32_t cur = st.cur;
    const long long seg_start = st.seg_start;
    if (j == 0) sFlag[gl] = flag;
    __syncthreads();

    // ---- in-block stitching, fixed order: chunk gl, gl+1, ...
    // role 0: this chunk's tail segment (starts here, continues);  role 1 (chunk 0 only): the
    // block's head segment (started in an earlier block).
    for (int role = 0; role < 2; ++role) {
        const bool is_tail = role == 0;
        if (is_tail ? !(flag & 4) : !(gl == 0 && (flag & 1))) continue;
        VecF<V> acc = vzero<V>();
        const float* src = is_tail ? sT : sH;
        if (j < nlane) acc = *reinterpret_cast<const VecF<V>*>(src + gl * tdim + j * V);
        float acc1 = is_tail ? sT1[gl] : sH1[gl];
        float gs = is_tail ? sTg[gl] : sHg[gl];
        bool closed = !is_tail && !(flag & 2);
        if (!closed) {
            for (int g2 = gl + 1; g2 < gpb; ++g2) {
                const int f2 = sFlag[g2];
                if (j < nlane) {
                    const VecF<V> h = *reinterpret_cast<const VecF<V>*>(sH + g2 * tdim + j * V);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc.v[v] += h.v[v];
                }
                acc1 += sH1[g2];
                gs += sHg[g2];
                if (!(f2 & 2)) { closed = true; break; }
            }
        }
        if (is_tail && closed) {
            write_row<V>(P, GR, a, coef, t_field[f], cur, f, seg_start, j, acc, acc1, gs);
        } else {
            // open at block level: the segment started in an earlier block (head) or leaves this
            // block (tail, not closed); the unit stitch pass finishes it
            float* dst2 = is_tail ? a.tail2 : a.head2;
            if (j < nlane) vstore<V>(dst2 + (size_t)blockIdx.x * tdim + j * V, acc);
            if (j == 0) {
                (is_tail ? a.tail1 : a.head1)[blockIdx.x] = acc1;
                (is_tail ? a.tailg : a.headg)[blockIdx.x] = gs;
                if (is_tail) {
                    a.tail_start[blockIdx.x] = seg_start; a.tail_field[blockIdx.x] = f;
                    a.open_list[atomicAdd(a.open_count, 1u)] = blockIdx.x;
                }
            }
        }
    }
    // counters: one pair of integer atomics per block (order-independent, so still deterministic)
    __shared__ int s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (j == 0 && (n_valid | n_heads)) { atomicAdd(&s_cnt[0], n_valid); atomicAdd(&s_cnt[1], n_heads); }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x])
        atomicAdd(a.counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

// ---- segments that leave their unit (listed by segreduce in open_list: only the unit in which the segment
// STARTS is listed).  stitch_kernel: one lane group per listed segment finds the unit in which it ends (G units
// per step) and, when it spans at most LONG_SPAN units, adds their head partials in unit order.  Longer
// segments (hot rows of small tables) are queued for stitch_long_kernel: one block each, lane group g adds the
// partials of units u+1+g, u+1+g+gpb, ... and group 0 folds tail(u) + partial(0) + partial(1) + ... -- fixed
// orders both, whichever block or group runs them.
constexpr int LONG_SPAN = 24;

template <int V>
__device__ __forceinline__ void stitch_write(const DevPlan& P, const DevGrads& GR, const BwdArgs& a, float coef,
                                             long long u, uint32_t k, int j, const VecF<V>& acc, float acc1, float gs) {
    const int f = a.tail_field[u];
    const FieldDev& fd = P.f[f];
    FieldB fb;
    fb.w2 = fd.w2; fb.w1 = fd.w1; fb.row_base = (unsigned)fd.row_base; fb.dim = fd.dim;
    fb.flat_off = fd.flat_off; fb.aux_off = fd.aux_off; fb.flags = 0;
    write_row<V>(P, GR, a, coef, fb, k, f, a.tail_start[u], j, acc, acc1, gs);
}

template <int V>
__global__ void __launch_bounds__(256)
stitch_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ DevGrads GR,
              const __grid_constant__ BwdArgs a, int G, long long unit) {
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G;
    const int j = threadIdx.x - gl * G;
    const unsigned gmask = group_mask(G);
    const unsigned lane_base = (threadIdx.x & 31u) & ~(unsigned)(G - 1);
    const int tdim = P.max_tdim, nlane = tdim / V;
    const float coef = a.l2x2 * (a.l2_gscale ? __ldg(a.l2_gscale) : 1.f);
    const unsigned n_open = *a.open_count;
    for (unsigned it = blockIdx.x * gpb + gl; it < n_open; it += gridDim.x * gpb) {
        const long long u = a.open_list[it];
        const uint32_t k = __ldg(a.skeys + (u + 1) * unit - 1);
        // last unit of the segment: the first unit after u that is not "through"
        long long last = u + 1;
        for (long long base = u + 1;; base += G) {
            const long long uu = base + j;
            const long long e = (uu + 1) * unit;
            const bool through = e < a.N && __ldg(a.skeys + e - 1) == k && __ldg(a.skeys + e) == k;
            unsigned stop = __ballot_sync(gmask, !through) >> lane_base;
            if (G < 32) stop &= (1u << G) - 1u;
            if (stop) { last = base + (__ffs(stop) - 1); break; }
        }
        if (last - u > LONG_SPAN) {
            if (j == 0) {
                const unsigned q = atomicAdd(a.long_count, 1u);
                a.long_list[2 * q] = (unsigned)u; a.long_list[2 * q + 1] = (unsigned)last;
            }
            continue;
        }
        VecF<V> acc = vzero<V>();
        if (j < nlane) acc = vload<V>(a.tail2 + (size_t)u * tdim + j * V);
        float acc1 = a.tail1[u], gs = a.tailg[u];
#pragma unroll 4
        for (long long uu = u + 1; uu <= last; ++uu) {
            if (j < nlane) {
                const VecF<V> h = vload<V>(a.head2 + (size_t)uu * tdim + j * V);
#pragma unroll
                for (int v = 0; v < V; ++v) acc.v[v] += h.v[v];
            }
            acc1 += __ldg(a.head1 + uu);
            gs += __ldg(a.headg + uu);
        }
        stitch_write<V>(P, GR, a, coef, u, k, j, acc, acc1, gs);
    }
}

template <int V>
__global__ void __launch_bounds__(256)
stitch_long_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ DevGrads GR,
                   const __grid_constant__ BwdArgs a, int G, long long unit) {
    extern __shared__ __align__(16) float st_sm[];
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G;
    const int j = threadIdx.x - gl * G;
    const int tdim = P.max_tdim, nlane = tdim / V;
    float* s_part = st_sm;                   // gpb x tdim
    float* s_p1 = s_part + gpb * tdim;       // gpb
    float* s_pg = s_p1 + gpb;                // gpb
    const float coef = a.l2x2 * (a.l2_gscale ? __ldg(a.l2_gscale) : 1.f);
    const unsigned n_long = *a.long_count;
    for (unsigned it = blockIdx.x; it < n_long; it += gridDim.x) {
        const long long u = a.long_list[2 * it], last = a.long_list[2 * it + 1];
        const uint32_t k = __ldg(a.skeys + (u + 1) * unit - 1);
        VecF<V> acc = vzero<V>();
        float acc1 = 0.f, gs = 0.f;
#pragma unroll 4
        for (long long uu = u + 1 + gl; uu <= last; uu += gpb) {
            if (j < nlane) {
                const VecF<V> h = vload<V>(a.head2 + (size_t)uu * tdim + j * V);
#pragma unroll
                for (int v = 0; v < V; ++v) acc.v[v] += h.v[v];
            }
            acc1 += __ldg(a.head1 + uu);
            gs += __ldg(a.headg + uu);
        }
        __syncthreads();                     // previous round's reads of s_part are done
        if (j < nlane) vstore<V>(s_part + gl * tdim + j * V, acc);
        if (j == 0) { s_p1[gl] = acc1; s_pg[gl] = gs; }
        __syncthreads();
        if (gl == 0) {
            VecF<V> tot = vzero<V>();
            if (j < nlane) tot = vload<V>(a.tail2 + (size_t)u * tdim + j * V);
            float t1 = a.tail1[u], tg = a.tailg[u];
            for (int g2 = 0; g2 < gpb; ++g2) {
                if (j < nlane) {
#pragma unroll
                    for (int v = 0; v < V; ++v) tot.v[v] += s_part[g2 * tdim + j * V + v];
                }
                t1 += s_p1[g2];
                tg += s_pg[g2];
            }
            stitch_write<V>(P, GR, a, coef, u, k, j, tot, t1, tg);
        }
    }
}

// ---- seg2: warp-sequential segmented reduction (the fast path of the plain shapes) ---------------------------
// Eligible plans (host check in embed_bwd_impl): every field has dim == fm_dim == D in {32, 64, 128}, id fields are
// plain SPARSE or sum / mean bags (the Criteo shapes, and the owner side of the sharded path).
// One WARP owns a span of `unit` consecutive sorted positions and walks it front to back; a lane owns VW = D / 32
// floats of the row, so every 128-, 256- or 512-byte gradient row is one coalesced warp load and the running
// segment sum lives in VW registers per lane.  Per tile of 32 positions each lane decodes ONE position (key, payload
// -> row offset, g_fm[b], first-order gradient, bag scale) and the warp then consumes the 32 positions in order with
// warp-uniform control flow: the decoded scalars travel by shuffle, NB positions' row loads are issued before the
// first is consumed, segment heads come from one ballot.  The table row of a segment (for -(sum g_fm) w + 2 l2 w) is
// requested when the segment starts and used when it ends.  No shared-memory staging, no block barrier in the loop,
// ~40 warp instructions per position (the lane-group kernel above: ~100).  Segments that cross span boundaries use
// the same head / tail carry records as segreduce_kernel, so stitch_kernel / stitch_long_kernel finish them
// unchanged.  Summation order inside a segment is the sorted (= batch) order: deterministic.
template <int VW> struct LaneVec;
template <> struct LaneVec<1> { using T = float; };
template <> struct LaneVec<2> { using T = float2; };
template <> struct LaneVec<4> { using T = float4; };

template <int VW>
__device__ __forceinline__ void lv_load_stream(float (&r)[VW], const float* p) {
    using T = typename LaneVec<VW>::T;
    const T t = __ldcs(reinterpret_cast<const T*>(p));
    const float* q = reinterpret_cast<const float*>(&t);
#pragma unroll
    for (int v = 0; v < VW; ++v) r[v] = q[v];
}
template <int VW>
__device__ __forceinline__ void lv_load(float (&r)[VW], const float* p) {
    using T = typename LaneVec<VW>::T;
    const T t = __ldg(reinterpret_cast<const T*>(p));
    const float* q = reinterpret_cast<const float*>(&t);
#pragma unroll
    for (int v = 0; v < VW; ++v) r[v] = q[v];
}
template <int VW>
__device__ __forceinline__ void lv_store(float* p, const float (&r)[VW]) {
    using T = typename LaneVec<VW>::T;
    T t;
    float* q = reinterpret_cast<float*>(&t);
#pragma unroll
    for (int v = 0; v < VW; ++v) q[v] = r[v];
    *reinterpret_cast<T*>(p) = t;
}

template <int VW>
struct Seg2State {
    uint32_t cur;          // key of the running segment
    int seg_start;         // sorted position of its first member
    int f;                 // its field
    bool lead;             // it started before this span: its sum goes to the head carry, not to a row
    float acc[VW], w[VW];  // running sum (this lane's floats); table row of the segment (requested at its head)
    float a1, gs, w1;
};

template <int VW>
__device__ __forceinline__ void seg2_close(const DevPlan& P, const DevGrads& GR, const BwdArgs& a, const FieldB* t_field,
                                           float coef, bool need_w, int lane, long long unit_idx, const Seg2State<VW>& st) {
    const int tdim = P.max_tdim;
    if (st.lead) {         // partial sum of a segment that started in an earlier span
        lv_store<VW>(a.head2 + (size_t)unit_idx * tdim + lane * VW, st.acc);
        if (lane == 0) { a.head1[unit_idx] = st.a1; a.headg[unit_idx] = st.gs; }
        return;
    }
    if (a.peer_n > 0) {
        float *pv, *ps;
        peer_dst(a, (long long)__ldg(a.uidx + st.seg_start), tdim, pv, ps);
        float o[VW];
#pragma unroll
        for (int v = 0; v < VW; ++v) o[v] = st.acc[v] * a.peer_scale;
        lv_store<VW>(pv + lane * VW, o);
        if (lane == 0) *reinterpret_cast<float4*>(ps) = make_float4(st.a1 * a.peer_scale, st.gs * a.peer_scale, 0.f, 0.f);
        return;
    }
    const FieldB& fb = t_field[st.f];
    const size_t row = (size_t)(st.cur - fb.row_base);
    float out[VW];
    const float cw = coef - st.gs;
#pragma unroll
    for (int v = 0; v < VW; ++v) out[v] = need_w ? fmaf(cw, st.w[v], st.acc[v]) : st.acc[v];
    float* dst = a.mode == DFM_GRAD_DENSE ? GR.g[st.f].gw2 + row * tdim : a.row_grad2 + (size_t)st.seg_start * tdim;
    lv_store<VW>(dst + lane * VW, out);
    if (lane == 0) {
        const float o1 = coef != 0.f ? fmaf(coef, st.w1, st.a1) : st.a1;
        if (a.mode == DFM_GRAD_DENSE) GR.g[st.f].gw1[row] = o1;
        else a.row_grad1[st.seg_start] = o1;
    }
}

template <int VW, bool HAS_FM, bool HAS_FIELD, bool HAS_BAG, bool DIRECT>
__global__ void __launch_bounds__(256, 4)
seg2_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ DevGrads GR,
            const __grid_constant__ BwdArgs a, long long unit) {
    __shared__ FieldB t_field[MAX_FIELDS];
    __shared__ unsigned short s_slotf[MAX_SLOTS];
    __shared__ int s_cnt[2];
    stage_fields(P, t_field);
    for (int s = threadIdx.x; s < P.S; s += blockDim.x) s_slotf[s] = P.slot_field[s];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    // positions whose row loads are in flight together: ~24 floats of load registers per lane
    constexpr int STREAMS = 1 + ((HAS_FM && !DIRECT) ? 1 : 0) + (HAS_FIELD ? 1 : 0) + (HAS_BAG ? 1 : 0);
    constexpr int NBQ = 24 / (VW * STREAMS);
    constexpr int NB = NBQ >= 8 ? 8 : NBQ >= 4 ? 4 : 2;
    const int lane = threadIdx.x & 31;
    const long long n_units = (a.N + unit - 1) / unit;
    const uint32_t PAD = a.pad_key;
    const int bits = a.slot_bits;
    const uint32_t smask = (1u << bits) - 1u;
    const float coef = a.peer_n > 0 ? 0.f : a.l2x2 * (a.l2_gscale ? __ldg(a.l2_gscale) : 1.f);
    // bag variants are compiled with HAS_FM and decide at run time; for the others HAS_FM says it all
    const bool fm_rt = HAS_FM && !DIRECT && (HAS_BAG ? a.g_fm != nullptr : true);
    const bool fm_on = DIRECT || fm_rt;
    const bool need_w = a.peer_n == 0 && (fm_on || coef != 0.f);
    const int tdim = P.max_tdim, D = P.D, T = P.T;
    // this lane's floats of row offset 0 of every stream (one IMAD.WIDE per row address in the loop)
    const float* __restrict__ gl = a.g_flat + lane * VW;
    const float* __restrict__ gfl = a.g_field + lane * VW;
    const float* __restrict__ sl = a.fm_sum + lane * VW;
    const float* __restrict__ fel = a.fe + lane * VW;
    int n_valid = 0, n_heads = 0;

    // Spans are claimed dynamically (one integer atomic per span): the sorted order groups the positions by table --
    // hot, L2-resident rows of the small tables first, cold rows of the big ones last -- so spans differ a lot in
    // cost; which warp takes which span does not change any result (every span has its own carry records).
    for (;;) {
        long long unit_idx = 0;
        if (lane == 0) unit_idx = (long long)atomicAdd(a.span_counter, 1u);
        unit_idx = __shfl_sync(0xffffffffu, unit_idx, 0);
        if (unit_idx >= n_units) break;
        const long long p_lo = unit_idx * unit;
        const long long p_hi = (p_lo + unit < a.N) ? p_lo + unit : a.N;
        Seg2State<VW> st;
        st.cur = PAD; st.seg_start = (int)p_lo; st.f = 0; st.lead = false; st.a1 = 0.f; st.gs = 0.f; st.w1 = 0.f;
#pragma unroll
        for (int v = 0; v < VW; ++v) { st.acc[v] = 0.f; st.w[v] = 0.f; }
        const uint32_t key0 = __ldg(a.skeys + p_lo);
        if (p_lo > 0 && key0 != PAD && __ldg(a.skeys + p_lo - 1) == key0) {     // the span opens inside a segment
            st.cur = key0; st.lead = true;
        }
        bool ended = false;
        for (long long p = p_lo; p < p_hi && !ended; p += 32) {
            // ---- decode: one position per lane
            const long long q = p + lane;
            const bool valid = q < p_hi;
            const uint32_t key = valid ? __ldg(a.skeys + q) : PAD;
            const uint32_t pay = valid ? __ldg(a.spay + q) : 0u;
            unsigned goff = 0, bofs = 0;     // float offsets of the position's gradient row / its sample's fm_sum row
            int fl = 0, bag = 0;
            float m = 0.f, o = 0.f, c = 1.f;
            if (key != PAD) {
                if (DIRECT) {
                    fl = field_of_key(t_field, P.n_fields, key);
                    if (a.g_sc) {                               // split layout: (M, tdim) vectors + (M, 4) scalars
                        goff = pay * (unsigned)tdim;
                        const float2 sc = __ldg(reinterpret_cast<const float2*>(a.g_sc) + 2 * (size_t)pay);
                        o = sc.x; m = sc.y;
                    } else {
                        goff = pay * (unsigned)a.row_stride;
                        o = __ldg(a.g_flat + goff + tdim);      // packed first-order gradient
                        m = __ldg(a.g_flat + goff + tdim + 1);  // packed g_fm (for -(sum g_fm) w)
                    }
                } else {
                    const uint32_t b = pay >> bits;
                    fl = s_slotf[pay & smask];
                    const FieldB& fb = t_field[fl];
                    goff = b * (unsigned)T + (unsigned)fb.flat_off;
                    bofs = b * (unsigned)D;
                    if (HAS_FM && a.g_fm) m = __ldg(a.g_fm + b);
                    if (a.g_first) o = __ldg(a.g_first + b);
                    if (HAS_BAG && (fb.flags & 0xf) == DFM_SEQUENCE) {
                        bag = 1;
                        if (((fb.flags >> 4) & 0xf) == DFM_MEAN) c = __uint_as_float(__ldg(a.aux + (size_t)b * P.A + fb.aux_off));
                        o *= c;
                    }
                }
            }
            const uint32_t prevk = __shfl_up_sync(0xffffffffu, key, 1);
            const unsigned headmask = __ballot_sync(0xffffffffu, key != (lane == 0 ? st.cur : prevk));
            const unsigned padmask = __ballot_sync(0xffffffffu, key == PAD);
            const int n_live = padmask ? __ffs(padmask) - 1 : 32;       // PAD keys sort last: everything after is PAD
            if (n_live < 32) ended = true;
            // ---- consume the tile in sorted order, NB positions per batch
            for (int r0 = 0; r0 < n_live; r0 += NB) {
                float gA[NB][VW], gB[HAS_FIELD ? NB : 1][VW], sv[HAS_FM ? NB : 1][VW], eB[HAS_BAG ? NB : 1][VW];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int r = (r0 + i) & 31;                       // positions >= n_live are loaded (valid memory) but not consumed
                    const unsigned go = __shfl_sync(0xffffffffu, goff, r);
                    lv_load_stream<VW>(gA[i], gl + go);
                    if (HAS_FIELD) lv_load_stream<VW>(gB[i], gfl + go);
                    if (HAS_FM && !DIRECT) {
                        const unsigned bo = __shfl_sync(0xffffffffu, bofs, r);
                        if (fm_rt) lv_load<VW>(sv[i], sl + bo);
                        else {
#pragma unroll
                            for (int v = 0; v < VW; ++v) sv[i][v] = 0.f;
                        }
                    }
                    if (HAS_BAG) {                                      // pooled embedding of the bag (aliased layout: same offset)
                        const int bg = __shfl_sync(0xffffffffu, bag, r);
                        if (bg && fm_rt) lv_load<VW>(eB[i], fel + go);
                        else {
#pragma unroll
                            for (int v = 0; v < VW; ++v) eB[i][v] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int r = r0 + i;
                    if (r >= n_live) break;
                    if ((headmask >> r) & 1u) {
                        if (st.cur != PAD) seg2_close<VW>(P, GR, a, t_field, coef, need_w, lane, unit_idx, st);
                        st.cur = __shfl_sync(0xffffffffu, key, r);
                        st.f = __shfl_sync(0xffffffffu, fl, r);
                        st.seg_start = (int)(p + r); st.lead = false; st.a1 = 0.f; st.gs = 0.f;
#pragma unroll
                        for (int v = 0; v < VW; ++v) st.acc[v] = 0.f;
                        ++n_heads;
                        if (need_w) {                                   // requested now, used when the segment ends
                            const FieldB& fb = t_field[st.f];
                            const size_t row = (size_t)(st.cur - fb.row_base);
                            lv_load<VW>(st.w, fb.w2 + row * tdim + lane * VW);
                            if (coef != 0.f) st.w1 = __ldg(fb.w1 + row);
                        }
                    }
                    const float mr = (HAS_FM || DIRECT) ? __shfl_sync(0xffffffffu, m, r) : 0.f;
                    const float orr = __shfl_sync(0xffffffffu, o, r);
                    float cr = 1.f;
                    int bg = 0;
                    if (HAS_BAG) { cr = __shfl_sync(0xffffffffu, c, r); bg = __shfl_sync(0xffffffffu, bag, r); }
#pragma unroll
                    for (int v = 0; v < VW; ++v) {
                        float t = gA[i][v];
                        if (HAS_FIELD) t += gB[i][v];
                        if (HAS_FM && !DIRECT) t = fmaf(mr, HAS_BAG ? sv[i][v] - eB[i][v] : sv[i][v], t);
                        if (HAS_BAG) t *= cr;
                        st.acc[v] += t;
                    }
                    if (!HAS_BAG || !bg) st.gs += mr;                   // bag members carry their e term themselves
                    st.a1 += orr;
                    ++n_valid;
                }
            }
        }
        // ---- the span's last segment: finished here, or handed to the stitch pass
        if (st.cur != PAD) {
            const bool continues = !ended && p_hi < a.N && __ldg(a.skeys + p_hi) == st.cur;
            if (st.lead || !continues) {
                seg2_close<VW>(P, GR, a, t_field, coef, need_w, lane, unit_idx, st);
            } else {
                lv_store<VW>(a.tail2 + (size_t)unit_idx * tdim + lane * VW, st.acc);
                if (lane == 0) {
                    a.tail1[unit_idx] = st.a1; a.tailg[unit_idx] = st.gs;
                    a.tail_start[unit_idx] = st.seg_start; a.tail_field[unit_idx] = st.f;
                    a.open_list[atomicAdd(a.open_count, 1u)] = (unsigned)unit_idx;
                }
            }
        }
    }
    // counters: one pair of integer atomics per block (order-independent, so still deterministic)
    if (lane == 0 && (n_valid | n_heads)) { atomicAdd(&s_cnt[0], n_valid); atomicAdd(&s_cnt[1], n_heads); }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x])
        atomicAdd(a.counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

template <int VW>
static void launch_seg2(bool direct, bool has_fm, bool has_field, bool has_bag, unsigned blocks, cudaStream_t st,
                        const DevPlan& P, const DevGrads& GR, const BwdArgs& a, long long unit) {
    if (direct) { seg2_kernel<VW, false, false, false, true><<<blocks, 256, 0, st>>>(P, GR, a, unit); return; }
    if (has_bag) {
        if (has_field) seg2_kernel<VW, true, true, true, false><<<blocks, 256, 0, st>>>(P, GR, a, unit);
        else seg2_kernel<VW, true, false, true, false><<<blocks, 256, 0, st>>>(P, GR, a, unit);
        return;
    }
    if (has_fm && has_field) seg2_kernel<VW, true, true, false, false><<<blocks, 256, 0, st>>>(P, GR, a, unit);
    else if (has_fm) seg2_kernel<VW, true, false, false, false><<<blocks, 256, 0, st>>>(P, GR, a, unit);
    else if (has_field) seg2_kernel<VW, false, true, false, false><<<blocks, 256, 0, st>>>(P, GR, a, unit);
    else seg2_kernel<VW, false, false, false, false><<<blocks, 256, 0, st>>>(P, GR, a, unit);
}

// ---- DENSE-field Linear grads and projection grads -------------------------------------
struct PgField {
    int f, nvals, part_off;   // part_off: offset (floats) of this field inside one slice's partials
};
// DENSE fields without projection are streamed (dense_stream_kernel); fields with a projection
// go through the shared-memory tile kernel (pgrads_kernel).  pf[] lists the tile fields first.
struct PgRun { short pf0, n; };   // streamed DENSE fields that are consecutive in the flat view
struct PgArgs {
    PgField pf[MAX_FIELDS];
    PgRun runs[MAX_FIELDS];
    int n_pf, n_tile, n_runs, n_slices, vals_per_slice;
    long long slice_len;
    float* partials;          // (n_slices, vals_per_slice)
};

// value layout of one field: [proj (D*d)] [gW2 (d)] [gb2 (d)] [gw1] [gb1]   (present parts only)
__global__ void __launch_bounds__(PG_THREADS)
pgrads_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ BwdArgs a,
              const __grid_constant__ PgArgs pg) {
    extern __shared__ float sm[];
    const PgField pfd = pg.pf[blockIdx.y];
    const int f = pfd.f;
    const FieldDev& fd = P.f[f];
    const int D = P.D, d = fd.dim, F = P.n_fields;
    const bool proj = fd.proj != nullptr, dense = fd.kind == DFM_DENSE;
    float* s_ge = sm;                         // PG_TILE x D
    float* s_gr = s_ge + PG_TILE * D;         // PG_TILE x d   g_raw (DENSE fields)
    float* s_raw = s_gr + PG_TILE * d;        // PG_TILE x d   raw row (projection)
    float* s_x = s_raw + PG_TILE * d;         // PG_TILE
    float* s_g1 = s_x + PG_TILE;              // PG_TILE
    float* s_acc = s_g1 + PG_TILE;            // nvals
    const int tid = threadIdx.x;
    for (int o = tid; o < pfd.nvals; o += PG_THREADS) s_acc[o] = 0.f;
    const long long b_lo = (long long)blockIdx.x * pg.slice_len;
    const long long b_hi = (b_lo + pg.slice_len < a.B) ? b_lo + pg.slice_len : a.B;
    const int n_proj = proj ? D * d : 0;
    for (long long t0 = b_lo; t0 < b_hi; t0 += PG_TILE) {
        const int ts = (int)((b_hi - t0 < PG_TILE) ? b_hi - t0 : PG_TILE);
        __syncthreads();
        for (int i = tid; i < PG_TILE * D; i += PG_THREADS) {
            const int s = i / D, k = i - s * D;
            float ge = 0.f;
            if (s < ts) {
                const long long b = t0 + s;
                const size_t eoff = ((size_t)b * F + f) * D + k;
                if (a.g_field) ge = __ldcs(a.g_field + eoff);
                if (a.g_fm) ge = fmaf(__ldg(a.g_fm + b), __ldg(a.fm_sum + (size_t)b * D + k) - __ldcs(a.fe + eoff), ge);
            }
            s_ge[i] = ge;
        }
        for (int i = tid; i < PG_TILE * d; i += PG_THREADS) {
            const int s = i / d, c = i - s * d;
            float gr = 0.f, raw = 0.f;
            if (s < ts) {
                const long long b = t0 + s;
                if (a.g_flat) gr = __ldcs(a.g_flat + (size_t)b * P.T + fd.flat_off + c);
                if (proj) raw = __ldcs(a.flat + (size_t)b * P.T + fd.flat_off + c);
            }
            s_gr[i] = gr;
            s_raw[i] = raw;
        }
        if (tid < PG_TILE) {
            const bool ok = tid < ts;
            s_x[tid] = (ok && dense) ? __ldg(reinterpret_cast<const float*>(fd.in) + t0 + tid) : 0.f;
            s_g1[tid] = (ok && a.g_first) ? __ldg(a.g_first + t0 + tid) : 0.f;
        }
        __syncthreads();
        if (dense) {   // g_raw = g_flat + P^T g_e   (or + g_e when there is no projection)
            for (int i = tid; i < PG_TILE * d; i += PG_THREADS) {
                const int s = i / d, c = i - s * d;
                float gr = s_gr[i];
                if (proj) {
                    for (int k = 0; k < D; ++k) gr = fmaf(s_ge[s * D + k], __ldg(fd.proj + (size_t)k * d + c), gr);
                } else {
                    gr += s_ge[s * D + c];
                }
                s_gr[i] = gr;
            }
            __syncthreads();
        }
        for (int o = tid; o < pfd.nvals; o += PG_THREADS) {
            float acc = 0.f;
            if (o < n_proj) {                       // dP[k][c] = sum_b g_e[b,k] * raw[b,c]
                const int k = o / d, c = o - k * d;
                for (int s = 0; s < PG_TILE; ++s) acc = fmaf(s_ge[s * D + k], s_raw[s * d + c], acc);
            } else {
                const int r = o - n_proj;
                if (r < d) { for (int s = 0; s < PG_TILE; ++s) acc = fmaf(s_gr[s * d + r], s_x[s], acc); }
                else if (r < 2 * d) { for (int s = 0; s < PG_TILE; ++s) acc += s_gr[s * d + (r - d)]; }
                else if (r == 2 * d) { for (int s = 0; s < PG_TILE; ++s) acc = fmaf(s_g1[s], s_x[s], acc); }
                else { for (int s = 0; s < PG_TILE; ++s) acc += s_g1[s]; }
            }
            s_acc[o] += acc;
        }
    }
    __syncthreads();
    float* out = pg.partials + (size_t)blockIdx.x * pg.vals_per_slice + pfd.part_off;
    for (int o = tid; o < pfd.nvals; o += PG_THREADS) out[o] = s_acc[o];
}

// DENSE fields without projection: g_raw = g_flat + g_field + g_fm (fm_sum - e) is streamed once.  One block
// owns a RUN of fields that are consecutive in the flat view and a slice of samples: thread (field u, lane c)
// reads dims [cV, cV+V) of field u for every sample of the slice -- the whole block reads one contiguous
// n*d*4-byte span per sample -- and keeps sum(g_raw * x), sum(g_raw) in registers, so there is no cross-thread
// reduction at all; the slices are added in order by pgrads_finish_kernel.
template <int V, bool HAS_FIELD>
__global__ void __launch_bounds__(PG_THREADS, 2)
dense_stream_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ BwdArgs a,
                    const __grid_constant__ PgArgs pg, int lanes) {
    const PgRun run = pg.runs[blockIdx.y];
    const int u = threadIdx.x / lanes, c = threadIdx.x - u * lanes;
    if (u >= run.n) return;
    const PgField pfd = pg.pf[pg.n_tile + run.pf0 + u];
    const int f = pfd.f;
    const FieldDev& fd = P.f[f];
    const int D = P.D, d = fd.dim, F = P.n_fields, nch = d / V;
    if (c >= nch) return;
    const long long b_lo = (long long)blockIdx.x * pg.slice_len;
    const long long b_hi = (b_lo + pg.slice_len < a.B) ? b_lo + pg.slice_len : a.B;
    VecF<V> aw = vzero<V>(), ab = vzero<V>();
    float a1w = 0.f, a1b = 0.f;
    // e[b, f, :] = x[b] * w2 + b2 is recomputed, not re-read
    const VecF<V> w2c = vload<V>(fd.w2 + c * V), b2c = vload<V>(fd.b2 + c * V);
    const float* xin = reinterpret_cast<const float*>(fd.in);
    const float* gp = a.g_flat ? a.g_flat + fd.flat_off + c * V : nullptr;
    const float* ep = HAS_FIELD ? a.g_field + (size_t)f * D + c * V : nullptr;
    const float* sp = a.g_fm ? a.fm_sum + c * V : nullptr;
    const bool first = c == 0 && a.g_first != nullptr;
    constexpr int UB = HAS_FIELD ? 4 : 8;
    for (long long b0 = b_lo; b0 < b_hi; b0 += UB) {
        VecF<V> g[UB], t[HAS_FIELD ? UB : 1], sv[UB];
        float x[UB], gfm[UB], g1[UB];
#pragma unroll
        for (int i = 0; i < UB; ++i) {       // every load of the batch is issued before any is consumed
            const long long b = (b0 + i < b_hi) ? b0 + i : b_hi - 1;
            x[i] = __ldg(xin + b);
            g[i] = gp ? vload_stream<V>(gp + (size_t)b * P.T) : vzero<V>();
            if (HAS_FIELD) t[i] = vload_stream<V>(ep + (size_t)b * F * D);
            sv[i] = sp ? vload<V>(sp + (size_t)b * D) : vzero<V>();
            gfm[i] = sp ? __ldg(a.g_fm + b) : 0.f;
            g1[i] = first ? __ldg(a.g_first + b) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < UB; ++i) {
            if (b0 + i >= b_hi) break;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                float gr = g[i].v[v];
                if (HAS_FIELD) gr += t[i].v[v];
                if (sp) gr = fmaf(gfm[i], sv[i].v[v] - fmaf(x[i], w2c.v[v], b2c.v[v]), gr);
                aw.v[v] = fmaf(gr, x[i], aw.v[v]);
                ab.v[v] += gr;
            }
            a1w = fmaf(g1[i], x[i], a1w);
            a1b += g1[i];
        }
    }
    float* out = pg.partials + (size_t)blockIdx.x * pg.vals_per_slice + pfd.part_off;
#pragma unroll
    for (int v = 0; v < V; ++v) { out[c * V + v] = aw.v[v]; out[d + c * V + v] = ab.v[v]; }
    if (c == 0) { out[2 * d] = a1w; out[2 * d + 1] = a1b; }
}

// Adds the per-slice partials in a fixed order: warp w of a block takes slices w, w + 8, ... for 32 consecutive
// values (coalesced), then warp 0 folds the 8 sums in order and adds the L2 term.
__global__ void __launch_bounds__(256)
pgrads_finish_kernel(const __grid_constant__ DevPlan P, const __grid_constant__ DevGrads GR,
                     const __grid_constant__ BwdArgs a, const __grid_constant__ PgArgs pg) {
    __shared__ float s_acc[8][32];
    const PgField pfd = pg.pf[blockIdx.y];
    const FieldDev& fd = P.f[pfd.f];
    const GradDev& gd = GR.g[pfd.f];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = blockIdx.x * 32 + lane;
    float acc = 0.f;
    if (o < pfd.nvals) {
        const float* src = pg.partials + pfd.part_off + o;
#pragma unroll 8
        for (int s = w; s < pg.n_slices; s += 8) acc += src[(size_t)s * pg.vals_per_slice];
    }
    s_acc[w][lane] = acc;
    __syncthreads();
    if (w != 0 || o >= pfd.nvals) return;
    acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc += s_acc[q][lane];
    const float coef = a.l2x2 * (a.l2_gscale ? __ldg(a.l2_gscale) : 1.f);
    const int d = fd.dim;
    const int n_proj = fd.proj ? P.D * d : 0;
    if (o < n_proj) { gd.gproj[o] = fmaf(coef, __ldg(fd.proj + o), acc); return; }
    const int r = o - n_proj;
    if (r < d) gd.gw2[r] = fmaf(coef, __ldg(fd.w2 + r), acc);
    else if (r < 2 * d) gd.gb2[r - d] = fmaf(coef, __ldg(fd.b2 + r - d), acc);
    else if (r == 2 * d) gd.gw1[0] = fmaf(coef, __ldg(fd.w1), acc);
    else gd.gb1[0] = fmaf(coef, __ldg(fd.b1), acc);
}

// ---- host-side workspace carving --------------------------------------------------------
struct BwdLayout {
    size_t cub_bytes, off_cub, off_payload, off_head2, off_head1, off_tail2, off_tail1, off_tstart, off_headg, off_tailg, off_tfield, off_counters,
        off_open, off_long, off_partials, total;
    long long n_chunks;
    int n_slices, vals_per_slice;
    long long slice_len;
};

static int sort_temp_bytes(long long n, int bits, size_t* out) {
    size_t bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, bits);
    if (e != cudaSuccess) { set_error("cub SortPairs size query: %s", cudaGetErrorString(e)); return DFM_ERR_CUDA; }
    *out = bytes;
    return DFM_OK;
}

static int pg_count_vals(const dfm_plan* plan, int f) {
    const int d = plan->dim[f];
    int n = 0;
    if (d != plan->fm_dim) n += plan->fm_dim * d;
    if (plan->kind[f] == DFM_DENSE) n += 2 * d + 2;
    return n;
}

static int make_layout(const dfm_plan* plan, long long B, BwdLayout& L, long long direct_rows = -1) {
    const long long N = direct_rows >= 0 ? direct_rows : B * plan->S;
    L.cub_bytes = 0;
    if (N > 0) { int rc = sort_temp_bytes(N, plan->key_bits, &L.cub_bytes); if (rc) return rc; }
    L.n_chunks = ceil_div(N > 0 ? N : 1, SEG2_UNIT);   // units: seg2 spans (segreduce blocks cover >= 256 positions)
    int vals = 0, n_pf = 0;
    for (int f = 0; f < plan->n_fields; ++f) { int n = pg_count_vals(plan, f); if (n) { vals += n; ++n_pf; } }
    L.vals_per_slice = vals;
    int want = n_pf ? 4 * sm_count() : 1;
    long long max_slices = ceil_div(B > 0 ? B : 1, 2 * PG_TILE);
    if (want > max_slices) want = (int)max_slices;
    if (want < 1) want = 1;
    L.slice_len = ceil_div(ceil_div(B > 0 ? B : 1, want), PG_TILE) * PG_TILE;
    L.n_slices = (int)ceil_div(B > 0 ? B : 1, L.slice_len);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.off_cub = take(L.cub_bytes);
    L.off_payload = take((size_t)N * 4);
    L.off_head2 = take((size_t)L.n_chunks * plan->max_tdim * 4);
    L.off_head1 = take((size_t)L.n_chunks * 4);
    L.off_tail2 = take((size_t)L.n_chunks * plan->max_tdim * 4);
    L.off_tail1 = take((size_t)L.n_chunks * 4);
    L.off_tstart = take((size_t)L.n_chunks * 8);
    L.off_headg = take((size_t)L.n_chunks * 4);
    L.off_tailg = take((size_t)L.n_chunks * 4);
    L.off_tfield = take((size_t)L.n_chunks * 4);
    L.off_counters = take(16);
    L.off_open = take((size_t)(L.n_chunks + 1) * 4);      // [0] = count, [1..] = owner units
    L.off_long = take((size_t)(L.n_chunks / 16 + 2) * 8); // [0] = count, then (owner, last) pairs of long segments
    L.off_partials = take((size_t)L.n_slices * vals * 4);
    L.total = off;
    return DFM_OK;
}

}  // namespace dfm

using namespace dfm;

extern "C" {

size_t dfm_embed_bwd_workspace_bytes(const dfm_plan* plan, int64_t batch) {
    if (!plan || batch < 0) return 0;
    BwdLayout L;
    if (make_layout(plan, batch, L) != DFM_OK) return 0;
    return L.total;
}

static int sort_keys_impl(const dfm_plan* plan, int64_t n, int S, const uint32_t* keys, uint32_t* sorted_keys,
                          uint32_t* sorted_payload, void* workspace, size_t workspace_bytes, void* stream);

size_t dfm_sort_keys_workspace_bytes(const dfm_plan* plan, int64_t n) {
    if (!plan || n <= 0) return 16;
    size_t cub_bytes = 0;
    if (sort_temp_bytes(n, plan->key_bits, &cub_bytes) != DFM_OK) return 0;
    return align_up(cub_bytes, 256) + (size_t)n * 4;
}

int dfm_sort_keys(const dfm_plan* plan, int64_t n, const uint32_t* keys, uint32_t* sorted_keys,
                  uint32_t* sorted_payload, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(plan, DFM_ERR_INVALID, "dfm_sort_keys: null argument");
    return sort_keys_impl(plan, n, plan->S, keys, sorted_keys, sorted_payload, workspace, workspace_bytes, stream);
}

// S = slots per sample of the payload encoding (1: payload = key position, the row-list mode)
static int sort_keys_impl(const dfm_plan* plan, int64_t n, int S, const uint32_t* keys, uint32_t* sorted_keys,
                          uint32_t* sorted_payload, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(plan && keys && sorted_keys && sorted_payload && workspace, DFM_ERR_INVALID, "dfm_sort_keys: null argument");
    if (n <= 0) return DFM_OK;
    DFM_REQUIRE(n < 0x7fffffffLL && ((n / S + 1) << slot_bits_of(S)) < 0xffffffffLL, DFM_ERR_UNSUPPORTED,
                "dfm_sort_keys: %lld keys do not fit a 32-bit sort", (long long)n);
    size_t cub_bytes = 0;
    int rc = sort_temp_bytes(n, plan->key_bits, &cub_bytes);
    if (rc) return rc;
    const size_t off_payload = align_up(cub_bytes, 256);
    DFM_REQUIRE(workspace_bytes >= off_payload + (size_t)n * 4, DFM_ERR_WORKSPACE,
                "dfm_sort_keys: workspace %zu < %zu", workspace_bytes, off_payload + (size_t)n * 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t* payload = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + off_payload);
    const int blocks = (int)(ceil_div(n, 256) < 8LL * sm_count() ? ceil_div(n, 256) : 8LL * sm_count());
    payload_kernel<<<blocks, 256, 0, st>>>(payload, n, S, slot_bits_of(S));
    DFM_CHECK_LAUNCH();
    DFM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(workspace, cub_bytes, keys, sorted_keys, payload, sorted_payload,
                                                   (int)n, 0, plan->key_bits, st));
    return DFM_OK;
}

static int embed_bwd_impl(const dfm_plan* plan, int64_t batch, long long direct_rows, const void* const* inputs,
                  const float* const* params, const float* g_first, const float* g_field,
                  const float* g_flat, const float* g_fm, const float* field_emb,
                  const float* flat, const float* fm_sum, const uint32_t* keys,
                  const uint32_t* aux, float l2, const float* l2_gscale, int mode,
                  float* const* grads, uint32_t* sorted_keys, uint32_t* sorted_payload,
                  float* row_grad2, float* row_grad1, int64_t* n_valid, void* workspace,
                  size_t workspace_bytes, void* stream) {
    const bool direct = direct_rows >= 0;
    const bool presorted = (mode & DFM_GRAD_PRESORTED) != 0;   // sorted_keys / sorted_payload already hold dfm_sort_keys' result
    mode &= ~DFM_GRAD_PRESORTED;
    const bool skip_tables = mode == DFM_GRAD_SKIP_TABLES;
    DFM_REQUIRE(plan && params && grads && workspace && (inputs || direct), DFM_ERR_INVALID, "dfm_embed_bwd: null argument");
    DFM_REQUIRE(batch >= 0, DFM_ERR_INVALID, "dfm_embed_bwd: negative batch");
    DFM_REQUIRE(mode == DFM_GRAD_DENSE || mode == DFM_GRAD_ROWSPARSE || (skip_tables && !direct), DFM_ERR_INVALID,
                "dfm_embed_bwd: unknown mode %d", mode);
    DFM_REQUIRE(!g_fm || (fm_sum && field_emb), DFM_ERR_INVALID, "dfm_embed_bwd: g_fm needs fm_sum and field_emb");
    DFM_REQUIRE(batch == 0 || plan->A == 0 || aux, DFM_ERR_INVALID, "dfm_embed_bwd: aux required");
    DFM_REQUIRE(batch == 0 || plan->S == 0 || skip_tables || (sorted_keys && sorted_payload && (keys || presorted)), DFM_ERR_INVALID, "dfm_embed_bwd: key buffers required");
    DFM_REQUIRE(batch == 0 || mode != DFM_GRAD_ROWSPARSE || plan->S == 0 || (row_grad2 && row_grad1 && n_valid), DFM_ERR_INVALID,
                "dfm_embed_bwd: row-sparse outputs required");
    DFM_REQUIRE(batch == 0 || plan->n_proj_expected == 0 || flat, DFM_ERR_INVALID, "dfm_embed_bwd: flat needed for projection grads");
    BwdLayout L;
    int rc = make_layout(plan, batch, L, direct_rows);
    if (rc) return rc;
    DFM_REQUIRE(workspace_bytes >= L.total, DFM_ERR_WORKSPACE, "dfm_embed_bwd: workspace %zu < %zu", workspace_bytes, L.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    const long long N = direct ? direct_rows : (long long)batch * plan->S;
    const int payS = direct ? 1 : plan->S;
    DFM_REQUIRE(N < 0x7fffffffLL && (direct || ((long long)batch << slot_bits_of(plan->S)) < 0xffffffffLL), DFM_ERR_UNSUPPORTED,
                "dfm_embed_bwd: batch * slots must fit 31 bits");

    DevPlan* P = new DevPlan;
    DevGrads* GR = new DevGrads;
    struct Guard { DevPlan* p; DevGrads* g; ~Guard() { delete p; delete g; } } guard{P, GR};
    int V = plan->fill(*P, inputs, params, true);
    P->aliased = field_emb == flat ? 1 : 0;
    if (direct) P->S = 0;      // no slot tables in the row-list mode (field comes from the key)
    memset(GR, 0, sizeof(*GR));
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    bool aligned = al16(g_flat) && al16(g_field) && al16(field_emb) && al16(fm_sum) && al16(row_grad2);
    for (int f = 0; f < plan->n_fields; ++f) {
        GradDev& g = GR->g[f];
        g.gw2 = grads[5 * f + 0]; g.gb2 = grads[5 * f + 1]; g.gw1 = grads[5 * f + 2];
        g.gb1 = grads[5 * f + 3]; g.gproj = grads[5 * f + 4];
        const bool table = plan->kind[f] != DFM_DENSE && !plan->foreign[f];
        if (table && mode == DFM_GRAD_DENSE) {
            DFM_REQUIRE(g.gw2 && g.gw1, DFM_ERR_INVALID, "dfm_embed_bwd: dense mode needs table grads for field %d", f);
            aligned = aligned && al16(g.gw2);
        }
        if (plan->kind[f] == DFM_DENSE && !direct) DFM_REQUIRE(g.gw2 && g.gb2 && g.gw1 && g.gb1, DFM_ERR_INVALID, "dfm_embed_bwd: DENSE field %d grads missing", f);
        if (plan->dim[f] != plan->fm_dim && !direct) DFM_REQUIRE(g.gproj, DFM_ERR_INVALID, "dfm_embed_bwd: projection grad of field %d missing", f);
    }
    if (!aligned) V = 1;
    const int lanes = plan->max_tdim > 0 ? plan->max_tdim / V : 1;
    DFM_REQUIRE(lanes <= 32, DFM_ERR_UNSUPPORTED, "dfm_embed_bwd: table dim %d needs %d lanes per row (max 32 x %d floats)",
                plan->max_tdim, lanes, V);
    const int G = next_pow2(lanes);

    BwdArgs a;
    memset(&a, 0, sizeof(a));
    a.g_first = g_first; a.g_field = g_field; a.g_flat = g_flat; a.g_fm = g_fm;
    a.fe = field_emb; a.flat = flat; a.fm_sum = fm_sum; a.aux = aux; a.l2_gscale = l2_gscale;
    a.l2x2 = 2.f * l2; a.mode = mode; a.B = batch; a.N = N;
    a.skeys = sorted_keys; a.spay = sorted_payload; a.row_grad2 = row_grad2; a.row_grad1 = row_grad1;
    a.head2 = reinterpret_cast<float*>(ws + L.off_head2); a.head1 = reinterpret_cast<float*>(ws + L.off_head1);
    a.tail2 = reinterpret_cast<float*>(ws + L.off_tail2); a.tail1 = reinterpret_cast<float*>(ws + L.off_tail1);
    a.tail_start = reinterpret_cast<long long*>(ws + L.off_tstart);
    a.headg = reinterpret_cast<float*>(ws + L.off_headg); a.tailg = reinterpret_cast<float*>(ws + L.off_tailg);
    a.slot_bits = slot_bits_of(payS);
    a.tail_field = reinterpret_cast<int*>(ws + L.off_tfield);
    a.direct = direct ? 1 : 0;
    a.row_stride = plan->max_tdim + 4;
    a.pad_key = (unsigned)plan->total_rows;
    if (direct && g_first) { a.g_sc = g_first; a.g_first = nullptr; }    // split layout of the exchanged rows (dfm_rows_bwd)
    a.counters = n_valid ? reinterpret_cast<unsigned long long*>(n_valid)
                         : reinterpret_cast<unsigned long long*>(ws + L.off_counters);
    a.open_count = reinterpret_cast<unsigned*>(ws + L.off_open);
    a.open_list = a.open_count + 1;
    a.long_count = reinterpret_cast<unsigned*>(ws + L.off_long);
    a.long_list = a.long_count + 2;
    a.span_counter = a.long_count + 1;
    const int fill_blocks = 8 * sm_count();

    // The DENSE-field / projection gradients (step 3) read only the upstream gradients and the inputs: they run on an
    // internal side stream forked HERE and joined at the end, underneath the sort / segmented reduction / stitch kernels
    // (the stitch kernels are latency-bound and leave the memory system idle; dense_stream is a pure HBM stream).
    // From the caller's point of view everything is still ordered on `stream`.
    cudaStream_t st3 = st;
    cudaEvent_t ev_join = nullptr;
    if (!direct && batch > 0 && N > 0 && !skip_tables && getenv("DFM_K2_FORK") == nullptr) {
        static cudaStream_t aux[16] = {};
        static cudaEvent_t ev_f[16] = {}, ev_j[16] = {};
        int dev = 0;
        DFM_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 16) {
            if (!aux[dev]) {
                DFM_CHECK_CUDA(cudaStreamCreateWithFlags(&aux[dev], cudaStreamNonBlocking));
                DFM_CHECK_CUDA(cudaEventCreateWithFlags(&ev_f[dev], cudaEventDisableTiming));
                DFM_CHECK_CUDA(cudaEventCreateWithFlags(&ev_j[dev], cudaEventDisableTiming));
            }
            DFM_CHECK_CUDA(cudaEventRecord(ev_f[dev], st));
            DFM_CHECK_CUDA(cudaStreamWaitEvent(aux[dev], ev_f[dev], 0));
            st3 = aux[dev];
            ev_join = ev_j[dev];
        }
    }
    struct Join {      // every exit path re-joins the side stream
        cudaStream_t main, side; cudaEvent_t ev;
        ~Join() { if (ev) { cudaEventRecord(ev, side); cudaStreamWaitEvent(main, ev, 0); } }
    } join{st, st3, ev_join};

    // 1. dense mode: every element of every table gradient starts as 2*l2*w (or 0)
    if (mode == DFM_GRAD_DENSE) {
        // small tensors (<= 1 M floats) are batched into one launch per 64; big tables keep their own full-width launch
        L2FillList* fl = new L2FillList;
        struct G3 { L2FillList* p; ~G3() { delete p; } } g3{fl};
        int nl = 0;
        auto flush = [&]() {
            if (nl > 0) l2_fill_multi_kernel<<<dim3(16, nl), 256, 0, st>>>(*fl, a.l2x2, l2_gscale);
            nl = 0;
        };
        auto add = [&](const float* p, float* g, long long n) {
            if (n > (1LL << 20)) {
                int b = (int)(ceil_div(n, 1024) < fill_blocks ? ceil_div(n, 1024) : fill_blocks);
                l2_fill_kernel<<<b, 256, 0, st>>>(p, g, n, a.l2x2, l2_gscale);
                return;
            }
            fl->p[nl] = p; fl->g[nl] = g; fl->n[nl] = n;
            if (++nl == 64) flush();
        };
        for (int f = 0; f < plan->n_fields; ++f) {
            if (plan->kind[f] == DFM_DENSE || plan->foreign[f]) continue;
            add(params[5 * f + 0], grads[5 * f + 0], plan->vocab[f] * plan->dim[f]);
            add(params[5 * f + 2], grads[5 * f + 2], plan->vocab[f]);
        }
        flush();
        DFM_CHECK_LAUNCH();
    }
    // 2. sort + segmented reduction of the id slots
    if (N > 0 && (batch > 0 || direct) && !skip_tables) {
        DFM_CHECK_CUDA(cudaMemsetAsync(a.counters, 0, 16, st));
        if (!presorted) {
            rc = sort_keys_impl(plan, N, payS, keys, sorted_keys, sorted_payload, ws + L.off_cub,
                                L.off_payload + (size_t)N * 4 - L.off_cub, stream);
            if (rc) return rc;
        }
        DFM_CHECK_CUDA(cudaMemsetAsync(a.open_count, 0, 4, st));
        DFM_CHECK_CUDA(cudaMemsetAsync(a.long_count, 0, 8, st));      // + the span counter next to it
        const int gpb = 256 / G;
        bool any_generic = false, any_generic_proj = false, any_bag = false;
        for (int f = 0; f < plan->n_fields; ++f) {
            any_generic = any_generic || plan->kind[f] == DFM_SEQUENCE || (plan->kind[f] == DFM_SPARSE && plan->dim[f] != plan->fm_dim);
            // what seg2 cannot do: a field (any kind) whose dim differs from fm_dim, a max-pooled bag
            any_generic_proj = any_generic_proj || plan->dim[f] != plan->fm_dim ||
                               (plan->kind[f] == DFM_SEQUENCE && plan->combiner[f] == DFM_MAX && !plan->foreign[f]);
            any_bag = any_bag || (plan->kind[f] == DFM_SEQUENCE && !plan->foreign[f]);
        }
        long long unit;
        // fast path (seg2): every field dim == fm_dim in {32, 64, 128}, id fields plain SPARSE or sum / mean bags
        bool fast = V == 4 && !any_generic_proj && plan->max_tdim == plan->fm_dim &&
                    (plan->fm_dim == 32 || plan->fm_dim == 64 || plan->fm_dim == 128) && (direct || (g_flat && plan->aliasable)) &&
                    (direct ? N * (long long)a.row_stride : (long long)batch * plan->T) < 0x7fffffffLL &&
                    getenv("DFM_K2_LEGACY") == nullptr;
        DFM_REQUIRE(fast || !(direct && a.g_sc), DFM_ERR_UNSUPPORTED,
                    "dfm_rows_bwd: the split (vector, scalar) row layout needs embedding_dim == fm_embed_dim in {32, 64, 128}");
        if (fast) {
            const int vw = plan->fm_dim / 32;
            unit = SEG2_UNIT;                                           // spans are claimed dynamically by the resident warps
            long long want_blocks = ceil_div(ceil_div(N, unit), 8);
            if (want_blocks > 4LL * sm_count()) want_blocks = 4LL * sm_count();   // 4 resident blocks of 8 warps per SM (64 registers)
            const unsigned blocks = (unsigned)want_blocks;
            const bool hf = g_fm != nullptr, hg = g_field != nullptr;
            if (vw == 1) launch_seg2<1>(direct, hf, hg, any_bag, blocks, st, *P, *GR, a, unit);
            else if (vw == 2) launch_seg2<2>(direct, hf, hg, any_bag, blocks, st, *P, *GR, a, unit);
            else launch_seg2<4>(direct, hf, hg, any_bag, blocks, st, *P, *GR, a, unit);
        } else {
            unit = (long long)gpb * CHUNK;               // positions per segreduce block
            const unsigned blocks = (unsigned)ceil_div(N, unit);
            const size_t smem = seg_smem_bytes(gpb, plan->max_tdim, (int)unit, direct ? 0 : plan->S, plan->n_fields) + 16;
            if (smem > 48 * 1024) {
                DFM_CHECK_CUDA(cudaFuncSetAttribute(segreduce_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                DFM_CHECK_CUDA(cudaFuncSetAttribute(segreduce_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                DFM_CHECK_CUDA(cudaFuncSetAttribute(segreduce_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                DFM_CHECK_CUDA(cudaFuncSetAttribute(segreduce_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
            if (V == 4) {
                if (any_generic) segreduce_kernel<4, true><<<blocks, 256, smem, st>>>(*P, *GR, a, G);
                else segreduce_kernel<4, false><<<blocks, 256, smem, st>>>(*P, *GR, a, G);
            } else {
                if (any_generic) segreduce_kernel<1, true><<<blocks, 256, smem, st>>>(*P, *GR, a, G);
                else segreduce_kernel<1, false><<<blocks, 256, smem, st>>>(*P, *GR, a, G);
            }
        }
        const long long n_units = ceil_div(N, unit);
        const long long want = ceil_div(n_units, gpb);
        const unsigned sblocks = (unsigned)(want < 4LL * sm_count() ? want : 4LL * sm_count());
        const unsigned lblocks = (unsigned)(n_units < 2LL * sm_count() ? n_units : 2LL * sm_count());
        const size_t ssm = (size_t)gpb * (plan->max_tdim + 2) * 4;
        if (V == 4) {
            stitch_kernel<4><<<sblocks, 256, 0, st>>>(*P, *GR, a, G, unit);
            stitch_long_kernel<4><<<lblocks, 256, ssm, st>>>(*P, *GR, a, G, unit);
        } else {
            stitch_kernel<1><<<sblocks, 256, 0, st>>>(*P, *GR, a, G, unit);
            stitch_long_kernel<1><<<lblocks, 256, ssm, st>>>(*P, *GR, a, G, unit);
        }
        DFM_CHECK_LAUNCH();
    } else if (n_valid) {
        DFM_CHECK_CUDA(cudaMemsetAsync(n_valid, 0, 16, st));
    }
    if (direct) return DFM_OK;   // the row-list mode has no DENSE-field / projection parameters
    // 3. DENSE-field Linear and projection gradients (tile fields first, then streamed fields)
    PgArgs* pg = new PgArgs;
    struct G2 { PgArgs* p; ~G2() { delete p; } } g2{pg};
    memset(pg, 0, sizeof(*pg));
    int max_smem_floats = 0, max_vals = 0, stream_lanes = 1;
    auto streamable = [&](int f) {
        return plan->kind[f] == DFM_DENSE && plan->dim[f] == plan->fm_dim && plan->dim[f] / V <= 32;
    };
    int last_stream_f = -2;
    for (int pass = 0; pass < 2; ++pass) {
        for (int f = 0; f < plan->n_fields; ++f) {
            const int n = pg_count_vals(plan, f);
            if (!n || (streamable(f) ? 1 : 0) != pass) continue;
            PgField& pf = pg->pf[pg->n_pf];
            pf.f = f; pf.nvals = n; pf.part_off = pg->vals_per_slice;
            pg->vals_per_slice += n; pg->n_pf++;
            if (n > max_vals) max_vals = n;
            if (pass == 0) {
                pg->n_tile++;
                const int fl = PG_TILE * (plan->fm_dim + 2 * plan->dim[f] + 2) + n;
                if (fl > max_smem_floats) max_smem_floats = fl;
            } else {
                stream_lanes = next_pow2(plan->dim[f] / V);   // every streamed field has dim == fm_dim
                // runs: consecutive fields (contiguous in the flat view), at most PG_THREADS / lanes per block
                const int si = pg->n_pf - 1 - pg->n_tile;
                if (pg->n_runs > 0 && f == last_stream_f + 1 && (pg->runs[pg->n_runs - 1].n + 1) * stream_lanes <= PG_THREADS) {
                    pg->runs[pg->n_runs - 1].n++;
                } else {
                    pg->runs[pg->n_runs].pf0 = (short)si; pg->runs[pg->n_runs].n = 1;
                    pg->n_runs++;
                }
                last_stream_f = f;
            }
        }
    }
    if (pg->n_pf > 0) {
        pg->n_slices = batch > 0 ? L.n_slices : 0; pg->slice_len = L.slice_len;
        pg->partials = reinterpret_cast<float*>(ws + L.off_partials);
        const size_t smem = (size_t)max_smem_floats * 4;
        DFM_REQUIRE(smem <= 200 * 1024, DFM_ERR_UNSUPPORTED, "dfm_embed_bwd: projection/dense grads need %zu B shared memory", smem);
        if (smem > 48 * 1024)
            DFM_CHECK_CUDA(cudaFuncSetAttribute(pgrads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (batch > 0 && pg->n_tile > 0)
            pgrads_kernel<<<dim3(pg->n_slices, pg->n_tile), PG_THREADS, smem, st3>>>(*P, a, *pg);
        if (batch > 0 && pg->n_runs > 0) {
            const dim3 grid(pg->n_slices, pg->n_runs);
            if (V == 4 && g_field) dense_stream_kernel<4, true><<<grid, PG_THREADS, 0, st3>>>(*P, a, *pg, stream_lanes);
            else if (V == 4) dense_stream_kernel<4, false><<<grid, PG_THREADS, 0, st3>>>(*P, a, *pg, stream_lanes);
            else if (g_field) dense_stream_kernel<1, true><<<grid, PG_THREADS, 0, st3>>>(*P, a, *pg, stream_lanes);
            else dense_stream_kernel<1, false><<<grid, PG_THREADS, 0, st3>>>(*P, a, *pg, stream_lanes);
        }
        pgrads_finish_kernel<<<dim3((unsigned)ceil_div(max_vals, 32), pg->n_pf), 256, 0, st3>>>(*P, *GR, a, *pg);
        DFM_CHECK_LAUNCH();
    }
    return DFM_OK;
}

int dfm_embed_bwd(const dfm_plan* plan, int64_t batch, const void* const* inputs,
                  const float* const* params, const float* g_first, const float* g_field,
                  const float* g_flat, const float* g_fm, const float* field_emb,
                  const float* flat, const float* fm_sum, const uint32_t* keys,
                  const uint32_t* aux, float l2, const float* l2_gscale, int mode,
                  float* const* grads, uint32_t* sorted_keys, uint32_t* sorted_payload,
                  float* row_grad2, float* row_grad1, int64_t* n_valid, void* workspace,
                  size_t workspace_bytes, void* stream) {
    return embed_bwd_impl(plan, batch, -1, inputs, params, g_first, g_field, g_flat, g_fm, field_emb, flat, fm_sum, keys,
                          aux, l2, l2_gscale, mode, grads, sorted_keys, sorted_payload, row_grad2, row_grad1, n_valid,
                          workspace, workspace_bytes, stream);
}

size_t dfm_sort_pairs_workspace_bytes(int64_t n, int bits) {
    if (n <= 0) return 16;
    size_t bytes = 0;
    if (sort_temp_bytes(n, bits, &bytes) != DFM_OK) return 0;
    return align_up(bytes, 256);
}

// stable LSB radix sort of (key, payload) pairs on the low `bits` bits of the key (CUB DeviceRadixSort)
int dfm_sort_pairs(int64_t n, int bits, const uint32_t* keys, const uint32_t* payload, uint32_t* sorted_keys,
                   uint32_t* sorted_payload, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(n >= 0 && bits >= 1 && bits <= 32, DFM_ERR_INVALID, "dfm_sort_pairs: bad argument");
    if (n == 0) return DFM_OK;
    DFM_REQUIRE(keys && payload && sorted_keys && sorted_payload && workspace && n < 0x7fffffffLL, DFM_ERR_INVALID, "dfm_sort_pairs: null tensor");
    size_t bytes = 0;
    int rc = sort_temp_bytes(n, bits, &bytes);
    if (rc) return rc;
    DFM_REQUIRE(workspace_bytes >= bytes, DFM_ERR_WORKSPACE, "dfm_sort_pairs: workspace %zu < %zu", workspace_bytes, bytes);
    DFM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(workspace, bytes, keys, sorted_keys, payload, sorted_payload, (int)n, 0, bits,
                                                   static_cast<cudaStream_t>(stream)));
    return DFM_OK;
}

size_t dfm_rows_bwd_workspace_bytes(const dfm_plan* plan, int64_t n_rows) {
    if (!plan || n_rows < 0) return 0;
    BwdLayout L;
    if (make_layout(plan, 0, L, n_rows) != DFM_OK) return 0;
    return L.total;
}

int dfm_rows_bwd(const dfm_plan* plan, int64_t n_rows, const float* const* params, const uint32_t* keys,
                 const float* g_rows, const float* g_scalars, float l2, const float* l2_gscale, int mode,
                 float* const* grads, uint32_t* sorted_keys, uint32_t* sorted_payload, float* row_grad2,
                 float* row_grad1, int64_t* n_valid, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(n_rows >= 0, DFM_ERR_INVALID, "dfm_rows_bwd: negative row count");
    DFM_REQUIRE(n_rows == 0 || ((keys || (mode & DFM_GRAD_PRESORTED)) && g_rows), DFM_ERR_INVALID, "dfm_rows_bwd: null argument");
    // split layout (g_scalars != NULL): g_rows is (n, tdim) and g_scalars (n, 4); it is carried in the g_first slot
    return embed_bwd_impl(plan, 0, n_rows, nullptr, params, g_scalars, nullptr, g_rows, nullptr, nullptr, nullptr, nullptr, keys,
                          nullptr, l2, l2_gscale, mode, grads, sorted_keys, sorted_payload, row_grad2, row_grad1, n_valid,
                          workspace, workspace_bytes, stream);
}

// Sample side of the row-sharded backward: segmented reduction of the batch's gradient rows per UNIQUE (owner, row)
// key, in the fixed sorted order, with the finished sums stored straight into the owners' exchange buffers.
int dfm_shard_bwd_peer(const dfm_plan* plan, int64_t batch, const float* g_first, const float* g_field, const float* g_flat,
                       const float* g_fm, const float* field_emb, const float* fm_sum, const uint32_t* aux,
                       const uint32_t* sorted_keys, const uint32_t* sorted_payload, const uint32_t* uidx, int64_t n_sorted,
                       uint32_t pad_key, int n_peers, const int64_t* peer_start, float* const* peer_vec, float* const* peer_sc,
                       float grad_scale, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(plan && batch >= 0 && n_sorted >= 0, DFM_ERR_INVALID, "dfm_shard_bwd_peer: bad argument");
    if (batch == 0 || n_sorted == 0) return DFM_OK;
    DFM_REQUIRE(g_flat && sorted_keys && sorted_payload && uidx && workspace && peer_start && peer_vec && peer_sc, DFM_ERR_INVALID,
                "dfm_shard_bwd_peer: null tensor");
    DFM_REQUIRE(n_peers >= 1 && n_peers <= 16, DFM_ERR_INVALID, "dfm_shard_bwd_peer: 1..16 peers");
    DFM_REQUIRE(!g_fm || (fm_sum && field_emb), DFM_ERR_INVALID, "dfm_shard_bwd_peer: g_fm needs fm_sum and field_emb");
    const int D = plan->fm_dim;
    bool ok = plan->aliasable && plan->max_tdim == D && (D == 32 || D == 64 || D == 128) && plan->vec == 4;
    bool any_bag = false;
    for (int f = 0; f < plan->n_fields; ++f) {
        ok = ok && plan->dim[f] == D;
        if (plan->kind[f] == DFM_SEQUENCE && plan->foreign[f]) {
            ok = ok && plan->combiner[f] != DFM_MAX;
            any_bag = true;
        }
    }
    DFM_REQUIRE(ok, DFM_ERR_UNSUPPORTED, "dfm_shard_bwd_peer: needs every embedding_dim == fm_embed_dim in {32, 64, 128} and sum / mean bags");
    DFM_REQUIRE((long long)batch * plan->T < 0x7fffffffLL && n_sorted < 0x7fffffffLL, DFM_ERR_UNSUPPORTED, "dfm_shard_bwd_peer: batch too large");
    DFM_REQUIRE(!any_bag || !g_fm || field_emb, DFM_ERR_INVALID, "dfm_shard_bwd_peer: bag fields need the field embeddings");
    DFM_REQUIRE(plan->A == 0 || aux, DFM_ERR_INVALID, "dfm_shard_bwd_peer: aux record required");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    DFM_REQUIRE(al16(g_flat) && al16(g_field) && al16(fm_sum) && al16(field_emb), DFM_ERR_UNSUPPORTED, "dfm_shard_bwd_peer: unaligned tensor");
    BwdLayout L;
    int rc = make_layout(plan, 0, L, n_sorted);
    if (rc) return rc;
    DFM_REQUIRE(workspace_bytes >= L.total, DFM_ERR_WORKSPACE, "dfm_shard_bwd_peer: workspace %zu < %zu", workspace_bytes, L.total);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    DevPlan* P = new DevPlan;
    DevGrads* GR = new DevGrads;
    struct Guard { DevPlan* p; DevGrads* g; ~Guard() { delete p; delete g; } } guard{P, GR};
    std::vector<const float*> dummy(5 * plan->n_fields, nullptr);
    plan->fill(*P, nullptr, dummy.data(), false);
    P->aliased = 1;
    memset(GR, 0, sizeof(*GR));
    BwdArgs a;
    memset(&a, 0, sizeof(a));
    a.g_first = g_first; a.g_field = g_field; a.g_flat = g_flat; a.g_fm = g_fm; a.fe = field_emb; a.fm_sum = fm_sum; a.aux = aux;
    a.mode = DFM_GRAD_ROWSPARSE; a.B = batch; a.N = n_sorted;
    a.skeys = sorted_keys; a.spay = sorted_payload;
    a.head2 = reinterpret_cast<float*>(ws + L.off_head2); a.head1 = reinterpret_cast<float*>(ws + L.off_head1);
    a.tail2 = reinterpret_cast<float*>(ws + L.off_tail2); a.tail1 = reinterpret_cast<float*>(ws + L.off_tail1);
    a.tail_start = reinterpret_cast<long long*>(ws + L.off_tstart);
    a.headg = reinterpret_cast<float*>(ws + L.off_headg); a.tailg = reinterpret_cast<float*>(ws + L.off_tailg);
    a.tail_field = reinterpret_cast<int*>(ws + L.off_tfield);
    a.slot_bits = slot_bits_of(plan->S);
    a.counters = reinterpret_cast<unsigned long long*>(ws + L.off_counters);
    a.open_count = reinterpret_cast<unsigned*>(ws + L.off_open); a.open_list = a.open_count + 1;
    a.long_count = reinterpret_cast<unsigned*>(ws + L.off_long); a.long_list = a.long_count + 2;
    a.span_counter = a.long_count + 1;
    a.pad_key = pad_key;
    a.peer_n = n_peers; a.uidx = uidx; a.peer_scale = grad_scale;
    for (int q = 0; q < n_peers; ++q) {
        DFM_REQUIRE(peer_vec[q] && peer_sc[q] && al16(peer_vec[q]) && al16(peer_sc[q]) && peer_start[q] <= peer_start[q + 1], DFM_ERR_INVALID,
                    "dfm_shard_bwd_peer: peer %d has a null / unaligned buffer or a negative range", q);
        a.peer_start[q] = peer_start[q]; a.peer_vec[q] = peer_vec[q]; a.peer_sc[q] = peer_sc[q];
    }
    a.peer_start[n_peers] = peer_start[n_peers];
    DFM_CHECK_CUDA(cudaMemsetAsync(a.counters, 0, 16, st));
    DFM_CHECK_CUDA(cudaMemsetAsync(a.open_count, 0, 4, st));
    DFM_CHECK_CUDA(cudaMemsetAsync(a.long_count, 0, 8, st));
    const long long N = n_sorted;
    const long long unit = SEG2_UNIT;
    long long want_blocks = ceil_div(ceil_div(N, unit), 8);
    if (want_blocks > 4LL * sm_count()) want_blocks = 4LL * sm_count();
    const unsigned blocks = (unsigned)want_blocks;
    const bool hf = g_fm != nullptr, hg = g_field != nullptr;
    const int vw = D / 32;
    if (vw == 1) launch_seg2<1>(false, hf, hg, any_bag, blocks, st, *P, *GR, a, unit);
    else if (vw == 2) launch_seg2<2>(false, hf, hg, any_bag, blocks, st, *P, *GR, a, unit);
    else launch_seg2<4>(false, hf, hg, any_bag, blocks, st, *P, *GR, a, unit);
    const int G = next_pow2(D / 4), gpb = 256 / G;
    const long long n_units = ceil_div(N, unit);
    const long long want = ceil_div(n_units, gpb);
    const unsigned sblocks = (unsigned)(want < 4LL * sm_count() ? want : 4LL * sm_count());
    const unsigned lblocks = (unsigned)(n_units < 2LL * sm_count() ? n_units : 2LL * sm_count());
    const size_t ssm = (size_t)gpb * (D + 2) * 4;
    stitch_kernel<4><<<sblocks, 256, 0, st>>>(*P, *GR, a, G, unit);
    stitch_long_kernel<4><<<lblocks, 256, ssm, st>>>(*P, *GR, a, G, unit);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
