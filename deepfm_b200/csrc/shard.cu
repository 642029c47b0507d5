// Row-sharded embedding tables across W GPUs:  owner(id) = id mod W,
// local_row = id div W, per field.  The exchange itself is an NCCL all-to-all issued by the host
// side (deepfm_b200/sharded.py); these kernels are the owner-side gather of the looked-up rows
// and the sample-side packing of the row gradients into send order.
//
//   shard_gather   keys (global rows, as received)  ->  vec (M, d), first-order (M), local keys (M)
//   shard_pack     upstream grads of the three views + FM  ->  g_vec (n, d), g_fo (n) in send order
#include <string.h>

#include "plan.cuh"

namespace dfm {

struct ShardField {
    const float* w2;       // local shard (rows_local, d)
    const float* w1;       // local shard (rows_local, 1)
    unsigned gbase;        // global row base of the field (key = gbase + id)
    unsigned lbase;        // local row base of the field on this rank
    int dim;
    int flat_off;          // sample side: offset of the field inside the flat view
    int field;             // sample side: field index
};
struct ShardArgs {
    ShardField f[MAX_FIELDS];
    int n;                 // number of table fields
    int world, rank, dmax;
    unsigned pad_local;    // local PAD key (= total local rows)
};

template <int V>
__global__ void __launch_bounds__(256)
shard_gather_kernel(const __grid_constant__ ShardArgs a, long long M, const uint32_t* __restrict__ keys, int G,
                    float* __restrict__ vec, float* __restrict__ fo, uint32_t* __restrict__ lkeys) {
    __shared__ ShardField t[MAX_FIELDS];
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) t[i] = a.f[i];
    __syncthreads();
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G, j = threadIdx.x - gl * G;
    for (long long i = (long long)blockIdx.x * gpb + gl; i < M; i += (long long)gridDim.x * gpb) {
        const uint32_t key = __ldg(keys + i);
        int fi = 0;
        for (int q = 1; q < a.n; ++q) if (key >= t[q].gbase) fi = q;
        const ShardField& sf = t[fi];
        const uint32_t id = key - sf.gbase;
        const uint32_t lrow = id / (uint32_t)a.world;         // id mod world == rank by construction
        if (j < sf.dim / V)
            vstore_stream<V>(vec + (size_t)i * a.dmax + j * V, vload<V>(sf.w2 + (size_t)lrow * sf.dim + j * V));
        if (j == 0) {
            fo[i] = __ldg(sf.w1 + lrow);
            lkeys[i] = id ? sf.lbase + lrow : a.pad_local;    // id 0 is the padding row: no gradient
        }
    }
}

struct PackArgs {
    const float *g_first, *g_field, *g_flat, *g_fm, *fe, *fm_sum;
    const long long* pos;  // (S, B): send position of slot s of sample b
    long long B;
    int S, T, F, D, dmax;
};

// one lane group per (slot, sample): g_vec[pos] = g_flat + g_field + g_fm (fm_sum - e);  g_fo[pos] = g_first
template <int V>
__global__ void __launch_bounds__(256)
shard_pack_kernel(const __grid_constant__ ShardArgs a, const __grid_constant__ PackArgs p, int G,
                  float* __restrict__ g_vec, float* __restrict__ g_fo) {
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G, j = threadIdx.x - gl * G;
    const long long n = p.B * p.S;
    for (long long i = (long long)blockIdx.x * gpb + gl; i < n; i += (long long)gridDim.x * gpb) {
        const long long b = i / p.S;
        const int s = (int)(i - b * p.S);
        const ShardField& sf = a.f[s];
        const long long q = __ldg(p.pos + (long long)s * p.B + b);
        if (j < sf.dim / V) {
            VecF<V> g = vzero<V>();
            if (p.g_flat) g = vload_stream<V>(p.g_flat + (size_t)b * p.T + sf.flat_off + j * V);
            const size_t eoff = ((size_t)b * p.F + sf.field) * p.D + j * V;
            if (p.g_field) {
                const VecF<V> t = vload_stream<V>(p.g_field + eoff);
#pragma unroll
                for (int v = 0; v < V; ++v) g.v[v] += t.v[v];
            }
            if (p.g_fm) {
                const float m = __ldg(p.g_fm + b);
                const VecF<V> sv = vload<V>(p.fm_sum + (size_t)b * p.D + j * V);
                const VecF<V> e = vload_stream<V>(p.fe + eoff);
#pragma unroll
                for (int v = 0; v < V; ++v) g.v[v] = fmaf(m, sv.v[v] - e.v[v], g.v[v]);
            }
            vstore_stream<V>(g_vec + (size_t)q * p.dmax + j * V, g);
        }
        if (j == 0) g_fo[q] = p.g_first ? __ldg(p.g_first + b) : 0.f;
    }
}

}  // namespace dfm

using namespace dfm;

static int fill_shard_args(const dfm_plan* local_plan, const int64_t* global_row_base, const float* const* params,
                           int world, int rank, ShardArgs& a, bool need_params) {
    memset(&a, 0, sizeof(a));
    a.world = world; a.rank = rank; a.dmax = local_plan->max_tdim; a.pad_local = (unsigned)local_plan->total_rows;
    int n = 0;
    for (int f = 0; f < local_plan->n_fields; ++f) {
        if (local_plan->kind[f] == DFM_DENSE) continue;
        if (local_plan->kind[f] != DFM_SPARSE || local_plan->dim[f] != local_plan->fm_dim) {
            set_error("sharded tables support SPARSE fields with embedding_dim == fm_embed_dim only (field %d)", f);
            return DFM_ERR_UNSUPPORTED;
        }
        ShardField& sf = a.f[n++];
        sf.w2 = need_params ? params[5 * f + 0] : nullptr;
        sf.w1 = need_params ? params[5 * f + 2] : nullptr;
        if (need_params && (!sf.w2 || !sf.w1)) { set_error("sharded: field %d table pointer is null", f); return DFM_ERR_INVALID; }
        sf.gbase = (unsigned)global_row_base[f];
        sf.lbase = (unsigned)local_plan->row_base[f];
        sf.dim = local_plan->dim[f];
        sf.flat_off = local_plan->flat_off[f];
        sf.field = f;
    }
    a.n = n;
    return DFM_OK;
}

extern "C" {

int dfm_shard_gather(const dfm_plan* local_plan, int world, int rank, const int64_t* global_row_base,
                     int64_t n_keys, const uint32_t* keys, const float* const* params, float* vec, float* fo,
                     uint32_t* local_keys, void* stream) {
    DFM_REQUIRE(local_plan && global_row_base && params && world > 0 && rank >= 0 && rank < world, DFM_ERR_INVALID,
                "dfm_shard_gather: bad argument");
    if (n_keys <= 0) return DFM_OK;
    DFM_REQUIRE(keys && vec && fo && local_keys, DFM_ERR_INVALID, "dfm_shard_gather: null tensor");
    ShardArgs* a = new ShardArgs;
    struct Gd { ShardArgs* p; ~Gd() { delete p; } } gd{a};
    int rc = fill_shard_args(local_plan, global_row_base, params, world, rank, *a, true);
    if (rc) return rc;
    const bool v4 = local_plan->vec == 4 && (reinterpret_cast<uintptr_t>(vec) & 15u) == 0;
    const int lanes = a->dmax / (v4 ? 4 : 1);
    DFM_REQUIRE(lanes <= 32, DFM_ERR_UNSUPPORTED, "dfm_shard_gather: table dim %d too wide", a->dmax);
    const int G = next_pow2(lanes);
    long long blocks = ceil_div(n_keys, 256 / G);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (v4) shard_gather_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*a, n_keys, keys, G, vec, fo, local_keys);
    else shard_gather_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*a, n_keys, keys, G, vec, fo, local_keys);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_shard_pack_grad(const dfm_plan* plan, int64_t batch, const int64_t* positions, const float* g_first,
                        const float* g_field, const float* g_flat, const float* g_fm, const float* field_emb,
                        const float* fm_sum, float* g_vec, float* g_fo, void* stream) {
    DFM_REQUIRE(plan && batch >= 0, DFM_ERR_INVALID, "dfm_shard_pack_grad: bad argument");
    if (batch == 0 || plan->S == 0) return DFM_OK;
    DFM_REQUIRE(positions && g_vec && g_fo && (!g_fm || (field_emb && fm_sum)), DFM_ERR_INVALID, "dfm_shard_pack_grad: null tensor");
    ShardArgs* a = new ShardArgs;
    struct Gd { ShardArgs* p; ~Gd() { delete p; } } gd{a};
    std::vector<int64_t> zeros(plan->n_fields + 1, 0);
    int rc = fill_shard_args(plan, zeros.data(), nullptr, 1, 0, *a, false);
    if (rc) return rc;
    DFM_REQUIRE(a->n == plan->S, DFM_ERR_UNSUPPORTED, "dfm_shard_pack_grad: one slot per table field expected");
    PackArgs p;
    p.g_first = g_first; p.g_field = g_field; p.g_flat = g_flat; p.g_fm = g_fm; p.fe = field_emb; p.fm_sum = fm_sum;
    p.pos = reinterpret_cast<const long long*>(positions); p.B = batch; p.S = plan->S; p.T = plan->T;
    p.F = plan->n_fields; p.D = plan->fm_dim; p.dmax = plan->max_tdim;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool v4 = plan->vec == 4 && al16(g_flat) && al16(g_field) && al16(field_emb) && al16(fm_sum) && al16(g_vec);
    const int lanes = p.dmax / (v4 ? 4 : 1);
    DFM_REQUIRE(lanes <= 32, DFM_ERR_UNSUPPORTED, "dfm_shard_pack_grad: table dim %d too wide", p.dmax);
    const int G = next_pow2(lanes);
    long long blocks = ceil_div(batch * plan->S, 256 / G);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (v4) shard_pack_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*a, p, G, g_vec, g_fo);
    else shard_pack_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*a, p, G, g_vec, g_fo);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
