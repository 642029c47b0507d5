// K1: FeatureEmbedding.forward fused with FMInteraction.forward.
//
// Reference call sites replaced (deepfm/models/layers/embedding.py:76-126, fm.py:18-23):
// one nn.Embedding / nn.EmbeddingBag / nn.Linear(1,d) per field and view, the optional
// projection Linear(d_f, D), torch.stack / sum / cat and the seven element-wise ops of the FM.
//
// Layout: a group of G = pow2(D / V) lanes owns one sample (V = 4 floats = one 128-bit access).
// Lane j owns dims [jV, jV+V) of every field embedding, so the per-dim FM accumulators
// (sum_f e, sum_f e^2) live in registers and the field embeddings are written to HBM exactly
// once and never read back.  The sample's ids are staged in shared memory first (one coalesced
// pass that also emits the sort keys of the backward and the SPARSE first-order terms), so every
// row address of the gather is known up front and the row loads of consecutive fields are
// issued back to back.
#include <stdarg.h>
#include <string.h>

#include "plan.cuh"

namespace dfm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------------------ forward kernel
template <int V>
__device__ __forceinline__ VecF<V> gather_row_chunk(const FieldDev& fd, int id, int c) {
    return vload<V>(fd.w2 + (size_t)id * fd.dim + c * V);
}

// Plain bag (SEQUENCE, dim == D, sum / mean, no projection): lane j owns dims [jV, jV+V) of the pooled
// row.  NBAG row gathers are in flight per lane (pads are predicated off, not branched around), added in
// bag order, so the summation order is the reference's (EmbeddingBag walks the bag front to back).
constexpr int NBAG = 8;
template <int V>
__device__ __forceinline__ VecF<V> pool_plain(const float* __restrict__ w2, const int* __restrict__ ids, int L,
                                              int rs, int j, int& cnt_out) {
    VecF<V> acc = vzero<V>();
    int cnt = 0;
    for (int l0 = 0; l0 < L; l0 += NBAG) {
        VecF<V> r[NBAG];
        int id[NBAG];
#pragma unroll
        for (int i = 0; i < NBAG; ++i) {
            id[i] = (l0 + i < L) ? ids[l0 + i] : 0;
            r[i] = vzero<V>();
            if (id[i]) r[i] = vload<V>(w2 + (size_t)id[i] * rs + j * V);
        }
#pragma unroll
        for (int i = 0; i < NBAG; ++i) {
            if (id[i]) {
                ++cnt;
#pragma unroll
                for (int v = 0; v < V; ++v) acc.v[v] += r[i].v[v];
            }
        }
    }
    cnt_out = cnt;
    return acc;
}

// Pool one chunk of a bag: EmbeddingBag(mode, padding_idx=0) -- pads skipped wherever they are,
// all-pad bag -> 0, duplicates count once per occurrence, max ties keep the first.
template <int V>
__device__ __forceinline__ VecF<V> pool_chunk(const FieldDev& fd, const int* s_ids, int c,
                                              float& inv_cnt, int* arg /* V entries or null */) {
    VecF<V> acc = vzero<V>();
    int cnt = 0;
    const int L = fd.max_len;
    if (fd.combiner == DFM_MAX) {
#pragma unroll
        for (int v = 0; v < V; ++v) { acc.v[v] = -INFINITY; arg[v] = 0; }
        for (int l = 0; l < L; ++l) {
            int id = s_ids[fd.slot_base + l];
            if (id == 0) continue;
            VecF<V> r = gather_row_chunk<V>(fd, id, c);
            ++cnt;
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (r.v[v] > acc.v[v]) { acc.v[v] = r.v[v]; arg[v] = l; }
        }
        if (cnt == 0) {
#pragma unroll
            for (int v = 0; v < V; ++v) { acc.v[v] = 0.f; arg[v] = -1; }
        }
        inv_cnt = 1.f;
        return acc;
    }
    for (int l = 0; l < L; ++l) {
        int id = s_ids[fd.slot_base + l];
        if (id == 0) continue;
        VecF<V> r = gather_row_chunk<V>(fd, id, c);
        ++cnt;
#pragma unroll
        for (int v = 0; v < V; ++v) acc.v[v] += r.v[v];
    }
    inv_cnt = 1.f;
    if (fd.combiner == DFM_MEAN) {
        inv_cnt = cnt > 0 ? 1.f / (float)cnt : 0.f;
        if (cnt > 0) {
            const float fc = (float)cnt;
#pragma unroll
            for (int v = 0; v < V; ++v) acc.v[v] = acc.v[v] / fc;  // true division like ATen
        }
    }
    return acc;
}

// Block-shared lookup tables, staged once per block from the by-value plan so that the per-slot /
// per-field loops read shared memory (lane-divergent constant-bank reads would serialise).
struct SlotS {              // one id slot
    const long long* in;    // id column
    const float* w1;        // first-order table
    unsigned row_base;      // key = row_base + id
    int vocab;
    int stride_pos;         // (ids per sample << 16) | position in the bag
    int sparse;             // 1: SPARSE (first-order term taken in phase 1); bit 1: foreign (no sort key)
    int w1s;                // floats between consecutive first-order weights
};
struct FieldS {             // one field, what the plain runs need
    const float* w2;
    const float* b2;
    int rs;                 // floats between consecutive table rows
};
struct DenseS {             // one DENSE field
    const float* in;
    const float* w1;
    const float* b1;
};

__host__ __device__ inline size_t fwd_table_bytes(int S, int F, int n_dense) {
    return (size_t)S * sizeof(SlotS) + (size_t)F * sizeof(FieldS) + (size_t)n_dense * sizeof(DenseS);
}

template <int V>
__global__ void __launch_bounds__(256, 3)
embed_fwd_kernel(const __grid_constant__ DevPlan P, long long B, int G, int smem_words_per_group,
                 float* __restrict__ first_order, float* __restrict__ field_emb,
                 float* __restrict__ flat, float* __restrict__ fm_out,
                 float* __restrict__ fm_sum, uint32_t* __restrict__ keys,
                 uint32_t* __restrict__ aux, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = P.D, F = P.n_fields, S = P.S, ND = P.n_dense;
    SlotS* t_slot = reinterpret_cast<SlotS*>(smem_raw);
    FieldS* t_field = reinterpret_cast<FieldS*>(t_slot + S);
    DenseS* t_dense = reinterpret_cast<DenseS*>(t_field + F);
    int* group_base = reinterpret_cast<int*>(smem_raw + ((fwd_table_bytes(S, F, ND) + 15) & ~(size_t)15));
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const FieldDev& fd = P.f[P.slot_field[s]];
        SlotS e;
        e.in = reinterpret_cast<const long long*>(fd.in); e.w1 = fd.w1;
        e.row_base = (unsigned)fd.row_base; e.vocab = fd.vocab;
        e.stride_pos = (fd.max_len << 16) | P.slot_pos[s];
        e.sparse = (fd.kind == DFM_SPARSE ? 1 : 0) | (fd.foreign ? 2 : 0);
        e.w1s = fd.w1_stride;
        t_slot[s] = e;
    }
    for (int f = threadIdx.x; f < F; f += blockDim.x) { t_field[f].w2 = P.f[f].w2; t_field[f].b2 = P.f[f].b2; t_field[f].rs = P.f[f].row_stride; }
    for (int i = threadIdx.x; i < ND; i += blockDim.x) {
        const FieldDev& fd = P.f[P.dense_field[i]];
        t_dense[i].in = reinterpret_cast<const float*>(fd.in); t_dense[i].w1 = fd.w1; t_dense[i].b1 = fd.b1;
    }
    __syncthreads();

    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G;
    const int j = threadIdx.x - gl * G;
    const unsigned gmask = group_mask(G);
    int* s_ids = group_base + gl * smem_words_per_group;
    float* s_x = reinterpret_cast<float*>(s_ids + S);
    float* s_raw = s_x + ND;
    const int nch_e = D / V;
    const long long n_tiles = (B + gpb - 1) / gpb;
    // persistent blocks: the tables above are staged once, then the block walks sample tiles
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long b = tile * gpb + gl;
    const bool active = b < B;
    const long long bb = active ? b : 0;          // inactive groups read sample 0, write nothing
    __syncwarp(gmask);                            // previous tile's reads of s_ids / s_x are done

    // ---- phase 1: ids and dense values -> shared memory; sort keys; SPARSE / DENSE first-order
    float fo_acc = 0.f;
    for (int s = j; s < S; s += G) {
        const SlotS e = t_slot[s];
        long long id = __ldg(e.in + bb * (e.stride_pos >> 16) + (e.stride_pos & 0xffff));
        if ((unsigned long long)id >= (unsigned long long)e.vocab) {   // also catches id < 0
            if (status) *status = 1;
            id = 0;
        }
        s_ids[s] = (int)id;
        if (keys && active) keys[b * S + s] = (id && !(e.sparse & 2)) ? e.row_base + (unsigned)id : P.pad_key;
        if (e.sparse & 1) fo_acc += __ldg(e.w1 + (size_t)id * e.w1s);  // row 0 returned as stored
    }
    for (int i = j; i < ND; i += G) {
        const DenseS e = t_dense[i];
        const float x = __ldg(e.in + bb);
        s_x[i] = x;
        fo_acc += x * __ldg(e.w1) + __ldg(e.b1);                      // Linear(1, 1)
    }
    __syncwarp(gmask);

    // ---- phase 2: gather / pool / affine, projection, FM accumulation
    VecF<V> Sacc = vzero<V>(), Qacc = vzero<V>();
    float* flat_b = flat + (size_t)bb * P.T;
    float* fe_b = field_emb + (size_t)bb * F * D;
    uint32_t* aux_b = aux ? aux + (size_t)bb * P.A : nullptr;
    const bool lane_on = j < nch_e;
    const bool two_views = !P.aliased;

    for (int ri = 0; ri < P.n_runs; ++ri) {
        const FieldRun run = P.runs[ri];
        if (run.cls == 0) {
            // plain SPARSE fields of dim D: row chunk -> flat (== field embedding), 4 gathers in flight
            if (lane_on) {
                const int* ids = s_ids + P.f[run.f0].slot_base;
                float* dst = flat_b + P.f[run.f0].flat_off + j * V;
                float* dst2 = fe_b + (size_t)run.f0 * D + j * V;
                const FieldS* tf = t_field + run.f0;
                int u = 0;
                // software pipeline: the next 4 row gathers are issued before the current 4 rows are
                // stored, so every lane always has gathers in flight
                VecF<V> r[4], nx[4];
                if (run.n >= 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) r[i] = vload<V>(tf[i].w2 + (size_t)ids[i] * tf[i].rs + j * V);
                }
                for (; u + 4 <= run.n; u += 4) {
                    const bool more = u + 8 <= run.n;
                    if (more) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) nx[i] = vload<V>(tf[u + 4 + i].w2 + (size_t)ids[u + 4 + i] * tf[u + 4 + i].rs + j * V);
                    }
                    if (active) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            vstore_stream<V>(dst + (u + i) * D, r[i]);
                            if (two_views) vstore_stream<V>(dst2 + (u + i) * D, r[i]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            Sacc.v[v] += r[i].v[v];
                            Qacc.v[v] += __fmul_rn(r[i].v[v], r[i].v[v]);
                        }
                    if (more) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) r[i] = nx[i];
                    }
                }
                for (; u < run.n; ++u) {
                    const VecF<V> r = vload<V>(tf[u].w2 + (size_t)ids[u] * tf[u].rs + j * V);
                    if (active) {
                        vstore_stream<V>(dst + u * D, r);
                        if (two_views) vstore_stream<V>(dst2 + u * D, r);
                    }
#pragma unroll
                    for (int v = 0; v < V; ++v) { Sacc.v[v] += r.v[v]; Qacc.v[v] += __fmul_rn(r.v[v], r.v[v]); }
                }
            }
            continue;
        }
        if (run.cls == 1) {
            // plain DENSE fields of dim D: Linear(1, D) on the staged scalar
            if (lane_on) {
                float* dst = flat_b + P.f[run.f0].flat_off + j * V;
                float* dst2 = fe_b + (size_t)run.f0 * D + j * V;
                const FieldS* tf = t_field + run.f0;
                for (int u = 0; u < run.n; ++u) {
                    const float x = s_x[run.x0 + u];
                    const VecF<V> w = vload<V>(tf[u].w2 + j * V);
                    const VecF<V> c = vload<V>(tf[u].b2 + j * V);
                    VecF<V> r;
#pragma unroll
                    for (int v = 0; v < V; ++v) r.v[v] = x * w.v[v] + c.v[v];
                    if (active) {
                        vstore_stream<V>(dst + u * D, r);
                        if (two_views) vstore_stream<V>(dst2 + u * D, r);
                    }
#pragma unroll
                    for (int v = 0; v < V; ++v) { Sacc.v[v] += r.v[v]; Qacc.v[v] += __fmul_rn(r.v[v], r.v[v]); }
                }
            }
            continue;
        }
        if (run.cls == 3) {
            // plain bags of dim D (sum / mean): pooled row -> flat (== field embedding)
            for (int u = 0; u < run.n; ++u) {
                const int f = run.f0 + u;
                const FieldDev& fd = P.f[f];
                const int* ids = s_ids + fd.slot_base;
                const int L = fd.max_len;
                const bool mean = fd.combiner == DFM_MEAN;
                int cnt = 0;
                if (lane_on) {
                    VecF<V> r = pool_plain<V>(fd.w2, ids, L, fd.row_stride, j, cnt);
                    if (mean && cnt > 0) {
                        const float fc = (float)cnt;
#pragma unroll
                        for (int v = 0; v < V; ++v) r.v[v] = r.v[v] / fc;   // true division like ATen
                    }
                    if (active) {
                        vstore_stream<V>(flat_b + fd.flat_off + j * V, r);
                        if (two_views) vstore_stream<V>(fe_b + (size_t)f * D + j * V, r);
                    }
#pragma unroll
                    for (int v = 0; v < V; ++v) { Sacc.v[v] += r.v[v]; Qacc.v[v] += __fmul_rn(r.v[v], r.v[v]); }
                }
                // first-order bag: one id per lane, fixed-order group reduction
                float w1sum = 0.f;
                int c1 = 0;
                for (int l = j; l < L; l += G) {
                    const int id = ids[l];
                    if (id) { w1sum += __ldg(fd.w1 + (size_t)id * fd.w1_stride); ++c1; }
                }
                w1sum = group_sum(w1sum, G, gmask);
                c1 = (int)group_sum((float)c1, G, gmask);
                if (j == 0) {
                    if (mean && c1 > 0) w1sum = w1sum / (float)c1;
                    fo_acc += w1sum;
                    if (mean && active && aux_b) aux_b[fd.aux_off] = __float_as_uint(c1 > 0 ? 1.f / (float)c1 : 0.f);
                }
            }
            continue;
        }
        // ---- generic field: sequence bags, projected fields (any kind)
        const int f = run.f0;
        const FieldDev& fd = P.f[f];
        const int nch = fd.dim / V;
        const bool proj = fd.proj != nullptr;
        float x = 0.f;
        int xi = 0;
        if (fd.kind == DFM_DENSE) {
            for (int i = 0; i < ND; ++i) if (P.dense_field[i] == f) xi = i;
            x = s_x[xi];
        }
        for (int c = j; c < nch; c += G) {
            VecF<V> r;
            if (fd.kind == DFM_SPARSE) {
                r = gather_row_chunk<V>(fd, s_ids[fd.slot_base], c);
            } else if (fd.kind == DFM_SEQUENCE) {
                float inv;
                int arg[V];
                r = pool_chunk<V>(fd, s_ids, c, inv, arg);
                if (active && aux_b) {
                    if (fd.combiner == DFM_MEAN && c == 0) aux_b[fd.aux_off] = __float_as_uint(inv);
                    if (fd.combiner == DFM_MAX) {
#pragma unroll
                        for (int v = 0; v < V; ++v) aux_b[fd.aux_off + c * V + v] = (uint32_t)arg[v];
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    r.v[v] = x * __ldg(fd.w2 + c * V + v) + __ldg(fd.b2 + c * V + v);  // Linear(1, d)
            }
            if (active) vstore_stream<V>(flat_b + fd.flat_off + c * V, r);
            if (!proj) {  // dim == D: chunk c of the raw row is chunk c of the field embedding
                if (active && two_views) vstore_stream<V>(fe_b + (size_t)f * D + c * V, r);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    Sacc.v[v] += r.v[v];
                    Qacc.v[v] += __fmul_rn(r.v[v], r.v[v]);
                }
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) s_raw[c * V + v] = r.v[v];
            }
        }
        if (proj) {  // e = raw @ P^T, P is (D, d) row-major (Linear(d, D, bias=False))
            __syncwarp(gmask);
            if (lane_on) {
                VecF<V> e = vzero<V>();
                const int d = fd.dim;
                for (int k = 0; k < d; ++k) {
                    const float rk = s_raw[k];
#pragma unroll
                    for (int v = 0; v < V; ++v) e.v[v] = fmaf(rk, __ldg(fd.proj + (size_t)(j * V + v) * d + k), e.v[v]);
                }
                if (active) vstore_stream<V>(fe_b + (size_t)f * D + j * V, e);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    Sacc.v[v] += e.v[v];
                    Qacc.v[v] += __fmul_rn(e.v[v], e.v[v]);
                }
            }
            __syncwarp(gmask);
        }
        // first-order term of a SEQUENCE field: one lane per field
        if (fd.kind == DFM_SEQUENCE && j == (f & (G - 1))) {
            float acc = fd.combiner == DFM_MAX ? -INFINITY : 0.f;
            int cnt = 0, arg = -1;
            for (int l = 0; l < fd.max_len; ++l) {
                int id = s_ids[fd.slot_base + l];
                if (id == 0) continue;
                float w = __ldg(fd.w1 + id);
                ++cnt;
                if (fd.combiner == DFM_MAX) {
                    if (w > acc) { acc = w; arg = l; }
                } else {
                    acc += w;
                }
            }
            if (cnt == 0) acc = 0.f;
            else if (fd.combiner == DFM_MEAN) acc = acc / (float)cnt;
            if (fd.combiner == DFM_MAX && active && aux_b) aux_b[fd.aux_off + fd.dim] = (uint32_t)arg;
            fo_acc += acc;
        }
    }

    // ---- epilogue: FM value 0.5 * sum_d (S_d^2 - Q_d), first-order sum, per-dim field sum
    float part = 0.f;
    if (lane_on) {
#pragma unroll
        for (int v = 0; v < V; ++v) part += __fmul_rn(Sacc.v[v], Sacc.v[v]) - Qacc.v[v];   // no contraction: one field => exactly 0
    }
    part = group_sum(part, G, gmask);
    fo_acc = group_sum(fo_acc, G, gmask);
    if (active) {
        if (j == 0) {
            first_order[b] = fo_acc;
            if (fm_out) fm_out[b] = 0.5f * part;
        }
        if (fm_sum && lane_on) vstore<V>(fm_sum + (size_t)b * D + j * V, Sacc);
    }
    }   // tile loop
}

// Keys only (bit-exact integer artefact; also the first stage of the input pipeline's ahead-of-step sort).
// The per-slot records are staged in shared memory (lane-divergent reads of the by-value plan would serialise
// in the constant bank); thread i owns key position i = b*S + s, so the key writes are coalesced.
__global__ void __launch_bounds__(256)
emit_keys_kernel(const __grid_constant__ DevPlan P, long long B, uint32_t* __restrict__ keys) {
    extern __shared__ __align__(16) unsigned char ek_raw[];
    SlotS* t_slot = reinterpret_cast<SlotS*>(ek_raw);
    const int S = P.S;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const FieldDev& fd = P.f[P.slot_field[s]];
        SlotS e;
        e.in = reinterpret_cast<const long long*>(fd.in); e.w1 = nullptr;
        e.row_base = (unsigned)fd.row_base; e.vocab = fd.vocab;
        e.stride_pos = (fd.max_len << 16) | P.slot_pos[s];
        e.sparse = fd.foreign ? 2 : 0;
        e.w1s = 0;
        t_slot[s] = e;
    }
    __syncthreads();
    const long long n = B * S;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / S;
        const int s = (int)(i - b * S);
        const SlotS e = t_slot[s];
        long long id = __ldg(e.in + b * (e.stride_pos >> 16) + (e.stride_pos & 0xffff));
        if ((unsigned long long)id >= (unsigned long long)e.vocab) id = 0;
        keys[i] = (id && !(e.sparse & 2)) ? e.row_base + (unsigned)id : P.pad_key;
    }
}

}  // namespace dfm

using namespace dfm;

// ------------------------------------------------------------------------------ plan (host)
int dfm_plan::fill(DevPlan& P, const void* const* inputs, const float* const* params,
                   bool need_inputs) const {
    memset(&P, 0, sizeof(P));
    P.n_fields = n_fields; P.D = fm_dim; P.T = T; P.S = S; P.A = A;
    P.aliased = 0; P.max_tdim = max_tdim; P.pad_key = (unsigned)total_rows;
    bool aligned = true;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    for (int f = 0; f < n_fields; ++f) {
        FieldDev& fd = P.f[f];
        fd.in = (inputs && need_inputs) ? inputs[f] : (inputs ? inputs[f] : nullptr);
        fd.w2 = params[5 * f + 0]; fd.b2 = params[5 * f + 1];
        fd.w1 = params[5 * f + 2]; fd.b1 = params[5 * f + 3];
        fd.proj = params[5 * f + 4];
        fd.row_base = row_base[f];
        fd.kind = kind[f]; fd.dim = dim[f]; fd.flat_off = flat_off[f]; fd.max_len = max_len[f];
        fd.combiner = combiner[f]; fd.slot_base = slot_base[f]; fd.aux_off = aux_off[f];
        fd.vocab = (int)vocab[f];
        fd.row_stride = row_stride[f] ? row_stride[f] : dim[f];
        fd.w1_stride = w1_stride[f] ? w1_stride[f] : 1;
        fd.foreign = foreign[f];
        if (kind[f] != DFM_DENSE) aligned = aligned && al16(fd.w2) && fd.row_stride % 4 == 0;
    }
    for (int s = 0; s < S; ++s) {
        P.slot_field[s] = (unsigned short)slot_field[s];
        P.slot_pos[s] = (unsigned short)slot_pos[s];
    }
    int n_dense = 0, n_runs = 0;
    for (int f = 0; f < n_fields; ++f) {
        const bool plain = dim[f] == fm_dim;
        const int cls = (plain && kind[f] == DFM_SPARSE) ? 0 : (plain && kind[f] == DFM_DENSE) ? 1
                        : (plain && kind[f] == DFM_SEQUENCE && combiner[f] != DFM_MAX) ? 3 : 2;
        if (n_runs > 0 && P.runs[n_runs - 1].cls == cls && cls != 2) {
            P.runs[n_runs - 1].n++;
        } else {
            P.runs[n_runs].cls = (short)cls; P.runs[n_runs].f0 = (short)f; P.runs[n_runs].n = 1;
            P.runs[n_runs].x0 = (short)n_dense;
            ++n_runs;
        }
        if (kind[f] == DFM_DENSE) P.dense_field[n_dense++] = (unsigned short)f;
    }
    P.n_runs = n_runs; P.n_dense = n_dense;
    return (vec == 4 && aligned) ? 4 : 1;
}

extern "C" {

const char* dfm_last_error(void) { return g_err; }
int dfm_version(void) { return 100; }

int dfm_device_info(int device, int64_t out[4]) {
    DFM_REQUIRE(out, DFM_ERR_INVALID, "dfm_device_info: out is null");
    cudaDeviceProp prop;
    DFM_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    out[0] = prop.multiProcessorCount; out[1] = prop.major; out[2] = prop.minor;
    out[3] = (int64_t)prop.sharedMemPerBlockOptin;
    return DFM_OK;
}

dfm_plan* dfm_plan_create(int n_fields, const int32_t* kind, const int32_t* dim,
                          const int64_t* vocab, const int32_t* max_len,
                          const int32_t* combiner, int fm_dim) {
    if (n_fields <= 0 || n_fields > MAX_FIELDS || !kind || !dim || !vocab || !max_len || !combiner || fm_dim <= 0) {
        set_error("dfm_plan_create: need 1..%d fields and non-null arrays (got %d)", MAX_FIELDS, n_fields);
        return nullptr;
    }
    dfm_plan* p = new dfm_plan();
    p->n_fields = n_fields; p->fm_dim = fm_dim;
    bool all4 = fm_dim % 4 == 0, alias = true;
    long long rows = 0;
    int T = 0, S = 0, A = 0, max_tdim = 0;
    for (int f = 0; f < n_fields; ++f) {
        const int k = kind[f], d = dim[f];
        const int L = k == DFM_SEQUENCE ? max_len[f] : 1;
        if (k < DFM_SPARSE || k > DFM_DENSE || d <= 0 || (k != DFM_DENSE && vocab[f] <= 0) || L <= 0 ||
            (k == DFM_SEQUENCE && (combiner[f] < DFM_SUM || combiner[f] > DFM_MAX)) || vocab[f] > 0x7fffffffLL) {
            set_error("dfm_plan_create: field %d has invalid kind/dim/vocab/max_len/combiner", f);
            delete p;
            return nullptr;
        }
        p->kind.push_back(k); p->dim.push_back(d); p->max_len.push_back(L);
        p->combiner.push_back(k == DFM_SEQUENCE ? combiner[f] : DFM_SUM);
        p->vocab.push_back(k == DFM_DENSE ? 0 : vocab[f]);
        p->flat_off.push_back(T); p->slot_base.push_back(S); p->aux_off.push_back(A);
        p->row_base.push_back(rows);
        p->row_stride.push_back(0); p->w1_stride.push_back(0); p->foreign.push_back(0);
        T += d;
        if (k != DFM_DENSE) {
            for (int l = 0; l < L; ++l) { p->slot_field.push_back(f); p->slot_pos.push_back(l); }
            S += L;
            rows += vocab[f];
            if (d > max_tdim) max_tdim = d;
        }
        if (k == DFM_SEQUENCE) A += combiner[f] == DFM_MEAN ? 1 : (combiner[f] == DFM_MAX ? d + 1 : 0);
        all4 = all4 && d % 4 == 0;
        alias = alias && d == fm_dim;
        if (d != fm_dim) p->n_proj_expected++;
    }
    p->row_base.push_back(rows);
    if (S > MAX_SLOTS || rows >= 0xffffffffLL) {
        set_error("dfm_plan_create: %d id slots (max %d) / %lld rows (max 2^32-2)", S, MAX_SLOTS, rows);
        delete p;
        return nullptr;
    }
    p->T = T; p->S = S; p->A = A; p->total_rows = rows; p->aliasable = alias ? 1 : 0;
    p->max_tdim = max_tdim; p->vec = all4 ? 4 : 1;
    int bits = 1;
    while ((1ULL << bits) <= (unsigned long long)rows) ++bits;   // PAD key == rows must sort last
    p->key_bits = bits;
    return p;
}

void dfm_plan_destroy(dfm_plan* plan) { delete plan; }

int dfm_plan_set_field_source(dfm_plan* plan, int field, int row_stride, int w1_stride, int foreign) {
    DFM_REQUIRE(plan && field >= 0 && field < plan->n_fields && row_stride >= 0 && w1_stride >= 0, DFM_ERR_INVALID,
                "dfm_plan_set_field_source: bad argument");
    const int f = field;
    DFM_REQUIRE(plan->kind[f] != DFM_DENSE, DFM_ERR_INVALID, "dfm_plan_set_field_source: field %d has no id table", f);
    if (row_stride || w1_stride)
        DFM_REQUIRE(plan->dim[f] == plan->fm_dim && (plan->kind[f] == DFM_SPARSE || plan->combiner[f] != DFM_MAX), DFM_ERR_UNSUPPORTED,
                    "dfm_plan_set_field_source: only plain SPARSE / sum- or mean-bag fields (dim == fm_dim) may use a strided row buffer");
    DFM_REQUIRE(row_stride == 0 || row_stride >= plan->dim[f], DFM_ERR_INVALID, "dfm_plan_set_field_source: stride < dim");
    plan->row_stride[f] = row_stride; plan->w1_stride[f] = w1_stride; plan->foreign[f] = foreign ? 1 : 0;
    plan->recompute_rows();
    if (plan->total_rows >= 0xffffffffLL) { set_error("dfm_plan_set_field_source: %lld rows (max 2^32-2)", plan->total_rows); return DFM_ERR_UNSUPPORTED; }
    return DFM_OK;
}

int dfm_plan_set_table_stride(dfm_plan* plan, int row_stride, int w1_stride) {
    DFM_REQUIRE(plan, DFM_ERR_INVALID, "dfm_plan_set_table_stride: null plan");
    for (int f = 0; f < plan->n_fields; ++f) {
        if (plan->kind[f] == DFM_DENSE) continue;
        int rc = dfm_plan_set_field_source(plan, f, row_stride, w1_stride, plan->foreign[f]);
        if (rc) return rc;
    }
    return DFM_OK;
}

int dfm_plan_info(const dfm_plan* plan, int64_t out[8]) {
    DFM_REQUIRE(plan && out, DFM_ERR_INVALID, "dfm_plan_info: null argument");
    out[0] = plan->T; out[1] = plan->S; out[2] = plan->total_rows; out[3] = plan->A;
    out[4] = plan->aliasable; out[5] = plan->max_tdim; out[6] = plan->key_bits; out[7] = plan->vec;
    return DFM_OK;
}

int dfm_plan_slots(const dfm_plan* plan, int32_t* slot_field, int32_t* slot_pos, int64_t* row_base) {
    DFM_REQUIRE(plan && slot_field && slot_pos && row_base, DFM_ERR_INVALID, "dfm_plan_slots: null argument");
    for (int s = 0; s < plan->S; ++s) { slot_field[s] = plan->slot_field[s]; slot_pos[s] = plan->slot_pos[s]; }
    for (int f = 0; f <= plan->n_fields; ++f) row_base[f] = plan->row_base[f];
    return DFM_OK;
}

int dfm_embed_fwd(const dfm_plan* plan, int64_t batch, const void* const* inputs,
                  const float* const* params, float* first_order, float* field_emb,
                  float* flat, float* fm_out, float* fm_sum, uint32_t* keys, uint32_t* aux,
                  int32_t* status, void* stream) {
    DFM_REQUIRE(batch >= 0, DFM_ERR_INVALID, "dfm_embed_fwd: negative batch");
    if (batch == 0 && plan) return DFM_OK;   // empty tensors have null data pointers
    DFM_REQUIRE(plan && inputs && params && first_order && field_emb && flat, DFM_ERR_INVALID,
                "dfm_embed_fwd: null argument");
    DFM_REQUIRE(plan->A == 0 || aux, DFM_ERR_INVALID, "dfm_embed_fwd: aux buffer required (%d words/sample)", plan->A);
    DFM_REQUIRE(field_emb != flat || plan->aliasable, DFM_ERR_INVALID,
                "dfm_embed_fwd: field_emb may alias flat only when every dim == fm_dim");
    DFM_REQUIRE((long long)batch * (plan->S > 0 ? plan->S : 1) < 0xffffffffLL, DFM_ERR_UNSUPPORTED,
                "dfm_embed_fwd: batch * slots must fit 32 bits");
    for (int f = 0; f < plan->n_fields; ++f) {
        DFM_REQUIRE(inputs[f] && params[5 * f] && params[5 * f + 2], DFM_ERR_INVALID, "dfm_embed_fwd: field %d has a null input/weight", f);
        DFM_REQUIRE(plan->kind[f] != DFM_DENSE || (params[5 * f + 1] && params[5 * f + 3]), DFM_ERR_INVALID,
                    "dfm_embed_fwd: DENSE field %d needs biases", f);
        DFM_REQUIRE((plan->dim[f] != plan->fm_dim) == (params[5 * f + 4] != nullptr), DFM_ERR_INVALID,
                    "dfm_embed_fwd: field %d projection must be present iff dim != fm_dim", f);
    }
    if (batch == 0) return DFM_OK;
    DevPlan* Pp = new DevPlan;   // 15 KB by-value kernel parameter: keep it off the stack
    int V = plan->fill(*Pp, inputs, params, true);
    Pp->aliased = field_emb == flat ? 1 : 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (!(al16(flat) && al16(field_emb) && (!fm_sum || al16(fm_sum)))) V = 1;
    const int nch_e = plan->fm_dim / V;
    if (nch_e > 32) {
        delete Pp;
        set_error("dfm_embed_fwd: fm_dim %d needs %d lanes per sample (max 32 x %d floats)", plan->fm_dim, nch_e, V);
        return DFM_ERR_UNSUPPORTED;
    }
    const int G = next_pow2(nch_e);
    int max_proj_dim = 0;
    for (int f = 0; f < plan->n_fields; ++f)
        if (plan->dim[f] != plan->fm_dim && plan->dim[f] > max_proj_dim) max_proj_dim = plan->dim[f];
    const int words = plan->S + Pp->n_dense + max_proj_dim;
    const int threads = 256;
    const int gpb = threads / G;
    const size_t smem = ((fwd_table_bytes(plan->S, plan->n_fields, Pp->n_dense) + 15) & ~(size_t)15) +
                        (size_t)gpb * words * sizeof(int);
    long long blocks = ceil_div(batch, gpb);
    if (blocks > 3LL * sm_count()) blocks = 3LL * sm_count();   // persistent: 3 resident blocks per SM
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    if (V == 4) {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(embed_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            embed_fwd_kernel<4><<<(unsigned)blocks, threads, smem, st>>>(*Pp, batch, G, words, first_order, field_emb, flat,
                                                                        fm_out, fm_sum, keys, aux, status);
    } else {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(embed_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            embed_fwd_kernel<1><<<(unsigned)blocks, threads, smem, st>>>(*Pp, batch, G, words, first_order, field_emb, flat,
                                                                        fm_out, fm_sum, keys, aux, status);
    }
    delete Pp;
    DFM_CHECK_CUDA(e);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_emit_keys(const dfm_plan* plan, int64_t batch, const void* const* inputs, uint32_t* keys, void* stream) {
    DFM_REQUIRE(plan && inputs && keys, DFM_ERR_INVALID, "dfm_emit_keys: null argument");
    if (batch <= 0 || plan->S == 0) return DFM_OK;
    std::vector<const float*> dummy(5 * plan->n_fields, nullptr);
    DevPlan* Pp = new DevPlan;
    plan->fill(*Pp, inputs, dummy.data(), true);
    const long long n = batch * plan->S;
    const int blocks = (int)(ceil_div(n, 256) < 8LL * sm_count() ? ceil_div(n, 256) : 8LL * sm_count());
    const size_t ek_smem = (size_t)plan->S * sizeof(SlotS);
    emit_keys_kernel<<<blocks, 256, ek_smem, static_cast<cudaStream_t>(stream)>>>(*Pp, batch, keys);
    delete Pp;
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
