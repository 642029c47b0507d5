// Row-sharded embedding tables across W GPUs:  owner(f, id) = (id + f) mod W,
// local_row = id div W, per field f (the rotation by the field index spreads the hot ids 1, 2, ... of the
// tables over the ranks).  The exchange itself is an NCCL all-to-all issued by the host
// side (deepfm_b200/sharded.py); these kernels are the owner-side gather of the looked-up rows
// and the sample-side packing of the row gradients into send order.
//
//   shard_route    id slots -> send order (stable grouping by owner), 1-based send positions per slot
//                  (0 = nothing sent: a padding entry of a multi-hot bag), per-owner counts
//   shard_gather   keys (global rows, as received)  ->  packed reply rows (M, d + 4), local keys (M)
//   shard_pack     upstream grads of the three views + FM  ->  packed gradient rows (n, d + 4) in send order
// Multi-hot bags (SEQUENCE fields, sum / mean) are exchanged id by id: the sample's GPU pools the
// received rows in K1 (global mean divisor = the bag's non-pad count) and the gradient of every member
// id travels back as its own row, so the owner-side reduction is the same sorted segmented sum.
#include <string.h>

#include "plan.cuh"

namespace dfm {

constexpr int RT_MAXW_ = 16;      // ranks (== RT_MAXW below)

struct ShardField {
    const float* w2;       // local shard (rows_local, d)
    const float* w1;       // local shard (rows_local, 1)
    unsigned gbase;        // global row base of the field (key = gbase + id)
    unsigned lbase;        // local row base of the field on this rank
    int dim;
    int flat_off;          // sample side: offset of the field inside the flat view
    int field;             // sample side: field index
    int slot_base, max_len;// sample side: first id slot / ids per sample of the field
    int bag;               // 0: SPARSE, 1: sum bag, 2: mean bag
    int aux_off;           // mean bag: word of the per-sample aux record that holds 1 / count
};
struct ShardArgs {
    ShardField f[MAX_FIELDS];
    int n;                 // number of table fields
    int world, rank, dmax;
    unsigned pad_local;    // local PAD key (= total local rows)
};

// Fused exchange over peer memory (NVLink P2P stores): row i of a rank's send order belongs to the peer whose
// segment [start[p], start[p + 1]) contains i and is written straight into that peer's buffer at
// base[p] + (i - start[p]) * row_stride -- the collective IS the kernel's store stream, there is no staging
// buffer and no separate all-to-all.  n == 0: plain local output.
struct PeerDst {
    long long start[RT_MAXW_ + 1];
    float* base[RT_MAXW_];
    int n;
};

__device__ __forceinline__ float* peer_row(const PeerDst& pd, float* local, long long i, int row_stride) {
    if (pd.n == 0) return local + (size_t)i * row_stride;
    int p = 0;
#pragma unroll 1
    for (int q = 1; q < pd.n; ++q) if (i >= pd.start[q]) p = q;
    return pd.base[p] + (size_t)(i - pd.start[p]) * row_stride;
}

template <int V>
__global__ void __launch_bounds__(256)
shard_gather_kernel(const __grid_constant__ ShardArgs a, long long M, const uint32_t* __restrict__ keys, int G,
                    float* __restrict__ vec, uint32_t* __restrict__ lkeys, const __grid_constant__ PeerDst pd) {
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
        const uint32_t lrow = id / (uint32_t)a.world;         // (id + field) mod world == rank by construction
        // packed reply row: [row (d), first-order weight, 0, 0, 0]
        float* dst = peer_row(pd, vec, i, a.dmax + 4);
        if (j < sf.dim / V)
            vstore_stream<V>(dst + j * V, vload<V>(sf.w2 + (size_t)lrow * sf.dim + j * V));
        if (j == 0) {
            *reinterpret_cast<float4*>(dst + a.dmax) = make_float4(__ldg(sf.w1 + lrow), 0.f, 0.f, 0.f);
            lkeys[i] = id ? sf.lbase + lrow : a.pad_local;    // id 0 is the padding row: no gradient
        }
    }
}

struct PackArgs {
    const float *g_first, *g_field, *g_flat, *g_fm, *fe, *fm_sum;
    const uint32_t* aux;
    const long long* pos;  // per field f: (B, max_len) block at B * slot_base[f]; 1-based send position, 0 = not sent
    long long B;
    int S, T, F, D, dmax, A;
    const uint32_t* send_slots;   // optional: send position -> id slot index b * S + s (the inverse of pos)
    long long n_sent;
    float grad_scale;             // multiplies every exchanged gradient row (1 / world: the global-mean-loss convention)
    unsigned short slot_tf[MAX_SLOTS];   // id slot -> index into ShardArgs::f
};

// one lane group per (sample, slot): packed gradient row in send order
//   SPARSE  [ g_flat + g_field + g_fm * fm_sum  (d) , g_first , g_fm , 0 , 0 ]
//           The owner folds -(sum g_fm) * w[row] in at the end of the segment (e == w[row] for every member),
//           so the field embeddings are not re-read here.
//   bag     [ scale * (g_flat + g_field + g_fm * (fm_sum - e_bag))  (d) , scale * g_first , 0 , 0 , 0 ]
//           scale = 1 (sum) or 1 / count (mean); e_bag is the pooled row K1 wrote.
template <int V>
__global__ void __launch_bounds__(256)
shard_pack_kernel(const __grid_constant__ ShardArgs a, const __grid_constant__ PackArgs p, int G,
                  float* __restrict__ g_vec, const __grid_constant__ PeerDst pd) {
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G, j = threadIdx.x - gl * G;
    // With send_slots the groups walk the SEND order: consecutive groups write consecutive rows of the same peer's
    // buffer (long sequential NVLink store streams; the random access moves to the local gradient reads).
    // Without it they walk the id slots and look the send position up.
    const long long n = p.send_slots ? p.n_sent : p.B * p.S;
    for (long long t = (long long)blockIdx.x * gpb + gl; t < n; t += (long long)gridDim.x * gpb) {
        const long long i = p.send_slots ? (long long)__ldg(p.send_slots + t) : t;
        const long long b = i / p.S;
        const int s = (int)(i - b * p.S);
        if (p.slot_tf[s] == 0xffff) continue;                  // replicated table: its gradient is local
        const ShardField& sf = a.f[p.slot_tf[s]];
        const long long q = p.send_slots ? t : __ldg(p.pos + p.B * sf.slot_base + b * sf.max_len + (s - sf.slot_base)) - 1;
        if (q < 0) continue;                                   // padding entry of a bag: nothing was sent
        float scale = 1.f;
        if (sf.bag == 2) scale = __uint_as_float(__ldg(p.aux + (size_t)b * p.A + sf.aux_off));
        const float m = p.g_fm ? __ldg(p.g_fm + b) : 0.f;
        const float gsc = p.grad_scale;
        float* dst = peer_row(pd, g_vec, q, p.dmax + 4);
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
                const VecF<V> sv = vload<V>(p.fm_sum + (size_t)b * p.D + j * V);
                if (sf.bag) {
                    const VecF<V> e = vload<V>(p.fe + eoff);
#pragma unroll
                    for (int v = 0; v < V; ++v) g.v[v] = fmaf(m, sv.v[v] - e.v[v], g.v[v]);
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v) g.v[v] = fmaf(m, sv.v[v], g.v[v]);
                }
            }
            if (sf.bag == 2 || gsc != 1.f) {
                const float sc = scale * gsc;
#pragma unroll
                for (int v = 0; v < V; ++v) g.v[v] *= sc;
            }
            vstore_stream<V>(dst + j * V, g);
        }
        if (j == 0)
            *reinterpret_cast<float4*>(dst + p.dmax) =
                make_float4((p.g_first ? __ldg(p.g_first + b) : 0.f) * scale * gsc, sf.bag ? 0.f : m * gsc, 0.f, 0.f);
    }
}

// ---- routing: stable grouping of the id slots by owner = id mod W (oracle.shard_route) --------
// count (per-block owner histogram) -> scan (one block) -> scatter (stable inside and across blocks).
constexpr int RT_TILE = 2048;     // id slots per block
constexpr int RT_MAXW = RT_MAXW_;  // ranks

struct RouteField {
    const long long* ids;   // (B,) or (B, max_len) id column
    unsigned gbase;         // global row base (key = gbase + id)
    int slot_base, max_len;
    int bag;                // SEQUENCE: padding entries (id 0) are not sent
    int rot;                // owner = (id + rot) mod W, rot = schema index of the field
    long long vocab;        // ids outside [0, vocab) are clamped to the padding id 0 and flagged (reference: IndexError)
};
struct RouteArgs {
    RouteField f[MAX_FIELDS];            // table fields only
    unsigned short slot_tf[MAX_SLOTS];   // id slot -> index into f
    int S, world;
    long long B;
    int* status;            // optional device word, set to 1 when an id is out of range
};

// owner of id slot i = b * S + s (sample-major source order), -1 if nothing is sent for it
__device__ __forceinline__ int route_owner(const RouteArgs& a, const RouteField* t, const unsigned short* slot_tf,
                                           long long i, long long& b, int& s, long long& id, const RouteField*& rf) {
    b = i / a.S; s = (int)(i - b * a.S);
    if (slot_tf[s] == 0xffff) { rf = nullptr; id = 0; return -1; }      // replicated table: looked up locally
    rf = t + slot_tf[s];
    id = __ldg(rf->ids + b * rf->max_len + (s - rf->slot_base));
    if ((unsigned long long)id >= (unsigned long long)rf->vocab) {     // also catches id < 0
        if (a.status) *a.status = 1;
        id = 0;
    }
    if (rf->bag && id == 0) return -1;
    return (int)((id + rf->rot) % a.world);
}

__global__ void __launch_bounds__(256)
route_count_kernel(const __grid_constant__ RouteArgs a, int* __restrict__ block_counts) {
    __shared__ int hist[RT_MAXW];
    __shared__ RouteField t[MAX_FIELDS];
    __shared__ unsigned short slot_tf[MAX_SLOTS];
    if (threadIdx.x < RT_MAXW) hist[threadIdx.x] = 0;
    for (int q = threadIdx.x; q < MAX_FIELDS; q += 256) t[q] = a.f[q];
    for (int q = threadIdx.x; q < a.S; q += 256) slot_tf[q] = a.slot_tf[q];
    __syncthreads();
    const long long n = a.B * a.S, i0 = (long long)blockIdx.x * RT_TILE;
    for (int c = threadIdx.x; c < RT_TILE; c += 256) {
        const long long i = i0 + c;
        if (i >= n) break;
        long long b, id; int s; const RouteField* rf;
        const int o = route_owner(a, t, slot_tf, i, b, s, id, rf);
        if (o >= 0) atomicAdd(&hist[o], 1);
    }
    __syncthreads();
    if (threadIdx.x < RT_MAXW) block_counts[blockIdx.x * RT_MAXW + threadIdx.x] = hist[threadIdx.x];
}

__global__ void route_scan_kernel(const int* __restrict__ block_counts, int nblk, int world,
                                  long long* __restrict__ offsets, long long* __restrict__ counts) {
    __shared__ long long total[RT_MAXW];
    const int w = threadIdx.x;
    if (w < world) {
        long long run = 0;
        for (int b = 0; b < nblk; ++b) { offsets[(size_t)w * nblk + b] = run; run += block_counts[b * RT_MAXW + w]; }
        total[w] = run;
        counts[w] = run;
    }
    __syncthreads();
    if (w < world) {
        long long base = 0;
        for (int q = 0; q < w; ++q) base += total[q];
        for (int b = 0; b < nblk; ++b) offsets[(size_t)w * nblk + b] += base;
    }
}

__global__ void __launch_bounds__(256)
route_scatter_kernel(const __grid_constant__ RouteArgs a, const long long* __restrict__ offsets, int nblk,
                     uint32_t* __restrict__ send_keys, long long* __restrict__ pos, uint32_t* __restrict__ send_slots) {
    __shared__ long long run[RT_MAXW];
    __shared__ int warp_cnt[8][RT_MAXW];
    __shared__ RouteField t[MAX_FIELDS];
    __shared__ unsigned short slot_tf[MAX_SLOTS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < a.world) run[threadIdx.x] = offsets[(size_t)threadIdx.x * nblk + blockIdx.x];
    for (int q = threadIdx.x; q < MAX_FIELDS; q += 256) t[q] = a.f[q];
    for (int q = threadIdx.x; q < a.S; q += 256) slot_tf[q] = a.slot_tf[q];
    const long long n = a.B * a.S, i0 = (long long)blockIdx.x * RT_TILE;
    for (int c = 0; c < RT_TILE; c += 256) {
        __syncthreads();
        const long long i = i0 + c + threadIdx.x;
        int o = -1, s = 0;
        long long b = 0, id = 0;
        const RouteField* rf = t;
        if (i < n) o = route_owner(a, t, slot_tf, i, b, s, id, rf);
        int rank = 0;
        for (int w = 0; w < a.world; ++w) {
            const unsigned m = __ballot_sync(0xffffffffu, o == w);
            if (o == w) rank = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) warp_cnt[warp][w] = __popc(m);
        }
        __syncthreads();
        if (i < n && rf) {
            long long q = -1;
            if (o >= 0) {
                q = run[o] + rank;
                for (int w2 = 0; w2 < warp; ++w2) q += warp_cnt[w2][o];
                send_keys[q] = rf->gbase + (uint32_t)id;
                if (send_slots) send_slots[q] = (uint32_t)i;
            }
            pos[a.B * rf->slot_base + b * rf->max_len + (s - rf->slot_base)] = q + 1;   // 1-based, 0 = not sent
        }
        __syncthreads();
        if (threadIdx.x < a.world) {
            int add = 0;
            for (int w2 = 0; w2 < 8; ++w2) add += warp_cnt[w2][threadIdx.x];
            run[threadIdx.x] += add;
        }
    }
}


// =================================================================================================================
// v2 exchange: UNIQUE rows per owner.  Under CTR-like skew most id slots of a batch repeat a row that the same GPU
// already asked the same owner for, so the sample side sorts its sharded id slots by the owner-major key
//     key = owner << lbits | (vbase[f] + id div W)           (the low part IS the owner's local sort key)
// and sends every distinct key once: send order = sorted unique keys (grouped by owner, ascending inside),
// positions[slot] = 1 + index of the slot's key in that order (0: nothing sent), and in the backward the gradient
// rows of all slots that share a key are pre-reduced on the sender in the same sorted order (dfm_shard_bwd_peer)
// before ONE row per unique key crosses NVLink.
//   shard_ukeys    id slots -> (key, payload = b << pbits | plan slot); unsent slots get the PAD key; positions = 0
//   shard_unique   sorted keys -> unique keys in send order, per-owner counts, unique index per sorted position,
//                  positions (count -> scan -> emit, stable, no float anywhere)
//   shard_gather2  owner side: unique local keys -> vector rows + [w1, 0, 0, 0] scalars, stored into the requester's
//                  buffers (peer stores), and the backward keys (PAD for the padding row id 0)
struct UKeyField {
    const long long* ids;
    unsigned vbase;            // local key base of the field (sum of ceil(V / W) of the sharded fields before it)
    int slot_base, max_len, bag, rot;
    long long vocab;
};
struct UKeyArgs {
    UKeyField f[MAX_FIELDS];             // sharded table fields
    unsigned short ts_field[MAX_SLOTS];  // compact sharded slot -> index into f
    unsigned short ts_pos[MAX_SLOTS];    // compact sharded slot -> position inside the bag
    int S_sh, world, lbits, pbits;
    unsigned pad;
    long long B;
    int* status;
};

__global__ void __launch_bounds__(256)
shard_ukeys_kernel(const __grid_constant__ UKeyArgs a, uint32_t* __restrict__ keys, uint32_t* __restrict__ payload,
                   long long* __restrict__ pos) {
    __shared__ UKeyField t[MAX_FIELDS];
    __shared__ unsigned short tf[MAX_SLOTS], tp[MAX_SLOTS];
    for (int q = threadIdx.x; q < MAX_FIELDS; q += 256) t[q] = a.f[q];
    for (int q = threadIdx.x; q < a.S_sh; q += 256) { tf[q] = a.ts_field[q]; tp[q] = a.ts_pos[q]; }
    __syncthreads();
    const long long n = a.B * a.S_sh;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const long long b = j / a.S_sh;
        const int ts = (int)(j - b * a.S_sh);
        const UKeyField& fd = t[tf[ts]];
        const int l = tp[ts];
        long long id = __ldg(fd.ids + b * fd.max_len + l);
        if ((unsigned long long)id >= (unsigned long long)fd.vocab) {     // also catches id < 0
            if (a.status) *a.status = 1;
            id = 0;
        }
        const bool sent = !(fd.bag && id == 0);
        const unsigned owner = (unsigned)((id + fd.rot) % a.world);
        keys[j] = sent ? (owner << a.lbits) | (fd.vbase + (unsigned)(id / a.world)) : a.pad;
        payload[j] = ((uint32_t)b << a.pbits) | (uint32_t)(fd.slot_base + l);
        pos[a.B * fd.slot_base + b * fd.max_len + l] = 0;
    }
}

constexpr int UQ_TILE = 2048;

__global__ void __launch_bounds__(256)
uniq_count_kernel(const uint32_t* __restrict__ sk, long long n, unsigned pad, int lbits, int* __restrict__ blk_heads,
                  int* __restrict__ blk_owner) {
    __shared__ int hist[RT_MAXW + 1];
    if (threadIdx.x <= RT_MAXW) hist[threadIdx.x] = 0;
    __syncthreads();
    const long long i0 = (long long)blockIdx.x * UQ_TILE;
    for (int c = threadIdx.x; c < UQ_TILE; c += 256) {
        const long long p = i0 + c;
        if (p >= n) break;
        const uint32_t k = __ldg(sk + p);
        if (k == pad) continue;
        if (p == 0 || __ldg(sk + p - 1) != k) { atomicAdd(&hist[RT_MAXW], 1); atomicAdd(&hist[k >> lbits], 1); }
    }
    __syncthreads();
    if (threadIdx.x < RT_MAXW) blk_owner[blockIdx.x * RT_MAXW + threadIdx.x] = hist[threadIdx.x];
    if (threadIdx.x == 0) blk_heads[blockIdx.x] = hist[RT_MAXW];
}

__global__ void uniq_scan_kernel(const int* __restrict__ blk_heads, const int* __restrict__ blk_owner, int nblk, int world,
                                 long long* __restrict__ blk_base, long long* __restrict__ counts) {
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int b = 0; b < nblk; ++b) { blk_base[b] = run; run += blk_heads[b]; }
    }
    if (threadIdx.x < world) {
        long long c = 0;
        for (int b = 0; b < nblk; ++b) c += blk_owner[b * RT_MAXW + threadIdx.x];
        counts[threadIdx.x] = c;
    }
}

struct UniqArgs {
    unsigned short slot_sb[MAX_SLOTS];   // plan slot -> slot_base of its field
    unsigned short slot_ml[MAX_SLOTS];   // plan slot -> max_len of its field
    long long B;
    int pbits;
    unsigned pad, lmask;
};

__global__ void __launch_bounds__(256)
uniq_emit_kernel(const __grid_constant__ UniqArgs a, const uint32_t* __restrict__ sk, const uint32_t* __restrict__ sp,
                 long long n, const long long* __restrict__ blk_base, uint32_t* __restrict__ ukeys,
                 uint32_t* __restrict__ uidx, long long* __restrict__ pos) {
    __shared__ int warp_cnt[8];
    __shared__ long long run;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) run = blk_base[blockIdx.x];
    const long long i0 = (long long)blockIdx.x * UQ_TILE;
    const uint32_t pmask = (1u << a.pbits) - 1u;
    for (int c = 0; c < UQ_TILE; c += 256) {
        __syncthreads();
        const long long p = i0 + c + threadIdx.x;
        uint32_t k = a.pad;
        bool head = false;
        if (p < n) {
            k = __ldg(sk + p);
            head = k != a.pad && (p == 0 || __ldg(sk + p - 1) != k);
        }
        const unsigned m = __ballot_sync(0xffffffffu, head);
        const int incl = __popc(m & (0xffffffffu >> (31 - lane)));      // heads at lanes <= this one
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = 0;
        for (int w2 = 0; w2 < warp; ++w2) before += warp_cnt[w2];
        if (k != a.pad) {
            const long long u = run + before + incl - 1;     // ordinal of this position's segment among all heads
            if (head) ukeys[u] = k & a.lmask;
            uidx[p] = (uint32_t)u;
            const uint32_t pay = __ldg(sp + p);
            const long long b = pay >> a.pbits;
            const int s = (int)(pay & pmask);
            pos[a.B * a.slot_sb[s] + b * a.slot_ml[s] + (s - a.slot_sb[s])] = u + 1;      // 1-based, 0 = not sent
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int add = 0;
            for (int w2 = 0; w2 < 8; ++w2) add += warp_cnt[w2];
            run += add;
        }
    }
}

struct PeerDst2 {
    long long start[RT_MAXW_ + 1];
    float* vec[RT_MAXW_];
    float* sc[RT_MAXW_];
    int n;
};

template <int V>
__global__ void __launch_bounds__(256)
shard_gather2_kernel(const __grid_constant__ ShardArgs a, long long M, const uint32_t* __restrict__ lk, int G,
                     uint32_t* __restrict__ bkeys, const __grid_constant__ PeerDst2 pd) {
    __shared__ ShardField t[MAX_FIELDS];
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) t[i] = a.f[i];
    __syncthreads();
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G, j = threadIdx.x - gl * G;
    for (long long i = (long long)blockIdx.x * gpb + gl; i < M; i += (long long)gridDim.x * gpb) {
        const uint32_t key = __ldg(lk + i);
        int fi = 0;
        for (int q = 1; q < a.n; ++q) if (key >= t[q].lbase) fi = q;
        const ShardField& sf = t[fi];
        const uint32_t lrow = key - sf.lbase;
        int p = 0;
#pragma unroll 1
        for (int q = 1; q < pd.n; ++q) if (i >= pd.start[q]) p = q;
        const long long r = i - pd.start[p];
        if (j < sf.dim / V)
            vstore_stream<V>(pd.vec[p] + (size_t)r * a.dmax + j * V, vload<V>(sf.w2 + (size_t)lrow * sf.dim + j * V));
        if (j == 0) {
            *reinterpret_cast<float4*>(pd.sc[p] + (size_t)r * 4) = make_float4(__ldg(sf.w1 + lrow), 0.f, 0.f, 0.f);
            // global id 0 (the padding row: local row 0 on rank f mod W) takes no gradient
            const bool id0 = lrow == 0 && ((a.rank - sf.field) % a.world + a.world) % a.world == 0;
            bkeys[i] = id0 ? a.pad_local : key;
        }
    }
}

}  // namespace dfm

using namespace dfm;

// Table fields of the plan that take part in the exchange: on the owner side (need_params) the fields whose
// gradient is produced here, on the sample side the FOREIGN fields (their tables live on the owners).
static bool shard_field(const dfm_plan* plan, int f, bool owner_side) {
    return plan->kind[f] != DFM_DENSE && (owner_side ? !plan->foreign[f] : plan->foreign[f] != 0);
}

static int fill_shard_args(const dfm_plan* local_plan, const int64_t* global_row_base, const float* const* params,
                           int world, int rank, ShardArgs& a, bool need_params) {
    memset(&a, 0, sizeof(a));
    a.world = world; a.rank = rank; a.dmax = local_plan->max_tdim; a.pad_local = (unsigned)local_plan->total_rows;
    int n = 0;
    for (int f = 0; f < local_plan->n_fields; ++f) {
        const int k = local_plan->kind[f];
        if (!shard_field(local_plan, f, need_params)) continue;
        if (local_plan->dim[f] != local_plan->fm_dim || local_plan->dim[f] % 4 ||
            (k == DFM_SEQUENCE && local_plan->combiner[f] == DFM_MAX)) {
            set_error("sharded tables support SPARSE and sum/mean SEQUENCE fields with embedding_dim == fm_embed_dim, "
                      "a multiple of 4 (field %d)", f);
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
        sf.slot_base = local_plan->slot_base[f];
        sf.max_len = local_plan->max_len[f];
        sf.bag = k == DFM_SEQUENCE ? (local_plan->combiner[f] == DFM_MEAN ? 2 : 1) : 0;
        sf.aux_off = local_plan->aux_off[f];
    }
    a.n = n;
    return DFM_OK;
}

// id slot -> index among the exchanged (sample side: foreign) table fields, 0xffff for the other slots
static void fill_slot_tf(const dfm_plan* plan, unsigned short* slot_tf) {
    std::vector<int> tf(plan->n_fields, 0xffff);
    int n = 0;
    for (int f = 0; f < plan->n_fields; ++f) if (shard_field(plan, f, false)) tf[f] = n++;
    for (int s = 0; s < plan->S; ++s) slot_tf[s] = (unsigned short)tf[plan->slot_field[s]];
}

static int fill_peer(PeerDst& pd, int n_peers, const int64_t* peer_start, float* const* peer_rows, const char* who) {
    memset(&pd, 0, sizeof(pd));
    if (n_peers == 0) return DFM_OK;
    if (n_peers < 0 || n_peers > RT_MAXW_ || !peer_start || !peer_rows) {
        set_error("%s: need 1..%d peers with segment starts and row pointers", who, RT_MAXW_);
        return DFM_ERR_INVALID;
    }
    for (int q = 0; q < n_peers; ++q) {
        if (!peer_rows[q] || (reinterpret_cast<uintptr_t>(peer_rows[q]) & 15u) || peer_start[q] > peer_start[q + 1]) {
            set_error("%s: peer %d has a null / unaligned row pointer or a negative segment", who, q);
            return DFM_ERR_INVALID;
        }
        pd.start[q] = peer_start[q]; pd.base[q] = peer_rows[q];
    }
    pd.start[n_peers] = peer_start[n_peers];
    pd.n = n_peers;
    return DFM_OK;
}

// All-ranks barrier on the stream, through peer-mapped flag words (one array of RT_MAXW_ words per rank, in the same
// symmetric allocation as the exchange buffers).  Lane r publishes `epoch` in word [rank] of rank r's array (system-scope
// release: every peer store this stream issued before is visible first), then waits until word [r] of its own array
// reaches `epoch` (acquire).  Epochs only grow (wrap-safe signed compare), so no reset pass is needed.  A peer that never
// arrives trips a ~4 s watchdog (trap) instead of hanging the GPU.
struct PeerFlags { uint32_t* flags[RT_MAXW_]; };
struct PeerI64 { long long* p[RT_MAXW_]; };
struct PeerU32 { uint32_t* p[RT_MAXW_]; };

// counts (W) of this rank -> row `rank` of the W x W count matrix of EVERY rank (peer-mapped stores)
__global__ void push_counts_kernel(const long long* __restrict__ counts, int world, int rank, PeerI64 dst) {
    const int t = threadIdx.x;
    if (t < world * world) dst.p[t / world][(size_t)rank * world + (t % world)] = counts[t % world];
}

// This rank's unique keys (send order: grouped by owner) -> the owners' receive buffers.  Offsets come from the count
// matrix M (device memory, complete after the barrier that follows push_counts): my segment for owner q starts at
// sum_{r<q} M[rank][r] in the send order and lands at sum_{s<rank} M[s][q] in q's receive order (grouped by source).
__global__ void __launch_bounds__(256)
push_keys_kernel(const uint32_t* __restrict__ send_keys, const long long* __restrict__ M, int world, int rank, long long cap,
                 PeerU32 dst) {
    __shared__ long long s_send[RT_MAXW_ + 1], s_recv[RT_MAXW_];
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int q = 0; q < world; ++q) { s_send[q] = acc; acc += M[(size_t)rank * world + q]; }
        s_send[world] = acc;
        for (int q = 0; q < world; ++q) {
            long long r = 0;
            for (int sr = 0; sr < rank; ++sr) r += M[(size_t)sr * world + q];
            s_recv[q] = r;
        }
    }
    __syncthreads();
    const long long n = s_send[world];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int q = 0;
        while (q + 1 < world && i >= s_send[q + 1]) ++q;
        const long long j = s_recv[q] + (i - s_send[q]);
        if (j < cap) dst.p[q][j] = send_keys[i];      // a receive buffer that would overflow is detected on the host (fits())
    }
}

__global__ void peer_barrier_kernel(PeerFlags pf, int world, int rank, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[r] + rank), "r"(epoch) : "memory");
    const uint32_t* mine = pf.flags[rank] + r;
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (clock64() - t0 > 8000000000LL) { printf("dfm peer barrier: rank %d timed out waiting for rank %d (epoch %u, saw %u)\n", rank, r, epoch, v); __trap(); }
        __nanosleep(64);
    }
    __threadfence_system();
}

extern "C" {

static int shard_gather_impl(const dfm_plan* local_plan, int world, int rank, const int64_t* global_row_base,
                             int64_t n_keys, const uint32_t* keys, const float* const* params, float* vec,
                             uint32_t* local_keys, int n_peers, const int64_t* peer_start, float* const* peer_rows,
                             void* stream) {
    DFM_REQUIRE(local_plan && global_row_base && params && world > 0 && rank >= 0 && rank < world, DFM_ERR_INVALID,
                "dfm_shard_gather: bad argument");
    if (n_keys <= 0) return DFM_OK;
    DFM_REQUIRE(keys && (vec || n_peers > 0) && local_keys, DFM_ERR_INVALID, "dfm_shard_gather: null tensor");
    ShardArgs* a = new ShardArgs;
    PeerDst* pd = new PeerDst;
    struct Gd { ShardArgs* p; PeerDst* q; ~Gd() { delete p; delete q; } } gd{a, pd};
    int rc = fill_shard_args(local_plan, global_row_base, params, world, rank, *a, true);
    if (rc) return rc;
    rc = fill_peer(*pd, n_peers, peer_start, peer_rows, "dfm_shard_gather_p2p");
    if (rc) return rc;
    DFM_REQUIRE(n_peers == 0 || peer_start[n_peers] - peer_start[0] == n_keys, DFM_ERR_INVALID,
                "dfm_shard_gather_p2p: the peer segments must cover the %lld keys", (long long)n_keys);
    const bool v4 = local_plan->vec == 4 && (reinterpret_cast<uintptr_t>(vec) & 15u) == 0;
    const int lanes = a->dmax / (v4 ? 4 : 1);
    DFM_REQUIRE(lanes <= 32, DFM_ERR_UNSUPPORTED, "dfm_shard_gather: table dim %d too wide", a->dmax);
    const int G = next_pow2(lanes);
    long long blocks = ceil_div(n_keys, 256 / G);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (v4) shard_gather_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*a, n_keys, keys, G, vec, local_keys, *pd);
    else shard_gather_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*a, n_keys, keys, G, vec, local_keys, *pd);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_shard_gather(const dfm_plan* local_plan, int world, int rank, const int64_t* global_row_base,
                     int64_t n_keys, const uint32_t* keys, const float* const* params, float* vec,
                     uint32_t* local_keys, void* stream) {
    return shard_gather_impl(local_plan, world, rank, global_row_base, n_keys, keys, params, vec, local_keys, 0, nullptr,
                             nullptr, stream);
}

int dfm_shard_gather_p2p(const dfm_plan* local_plan, int world, int rank, const int64_t* global_row_base,
                         int64_t n_keys, const uint32_t* keys, const float* const* params, int n_peers,
                         const int64_t* peer_start, float* const* peer_rows, uint32_t* local_keys, void* stream) {
    DFM_REQUIRE(n_peers > 0, DFM_ERR_INVALID, "dfm_shard_gather_p2p: no peers");
    return shard_gather_impl(local_plan, world, rank, global_row_base, n_keys, keys, params, nullptr, local_keys, n_peers,
                             peer_start, peer_rows, stream);
}

static int shard_pack_impl(const dfm_plan* plan, int64_t batch, const int64_t* positions, const float* g_first,
                           const float* g_field, const float* g_flat, const float* g_fm, const float* fm_sum,
                           const float* field_emb, const uint32_t* aux, float* g_vec, int n_peers,
                           const int64_t* peer_start, float* const* peer_rows, const uint32_t* send_slots, int64_t n_sent,
                           float grad_scale, void* stream) {
    DFM_REQUIRE(plan && batch >= 0, DFM_ERR_INVALID, "dfm_shard_pack_grad: bad argument");
    if (batch == 0 || plan->S == 0) return DFM_OK;
    DFM_REQUIRE(positions && (g_vec || n_peers > 0) && (!g_fm || fm_sum), DFM_ERR_INVALID, "dfm_shard_pack_grad: null tensor");
    ShardArgs* a = new ShardArgs;
    PackArgs* pp = new PackArgs;
    PeerDst* pd = new PeerDst;
    struct Gd { ShardArgs* p; PackArgs* q; PeerDst* r; ~Gd() { delete p; delete q; delete r; } } gd{a, pp, pd};
    std::vector<int64_t> zeros(plan->n_fields + 1, 0);
    int rc = fill_shard_args(plan, zeros.data(), nullptr, 1, 0, *a, false);
    if (rc) return rc;
    rc = fill_peer(*pd, n_peers, peer_start, peer_rows, "dfm_shard_pack_grad_p2p");
    if (rc) return rc;
    bool any_bag = false, any_mean = false;
    for (int i = 0; i < a->n; ++i) { any_bag = any_bag || a->f[i].bag; any_mean = any_mean || a->f[i].bag == 2; }
    DFM_REQUIRE(!any_bag || !g_fm || field_emb, DFM_ERR_INVALID, "dfm_shard_pack_grad: bag fields need the field embeddings for the FM gradient");
    DFM_REQUIRE(!any_mean || aux, DFM_ERR_INVALID, "dfm_shard_pack_grad: mean bags need the aux record of the forward");
    PackArgs& p = *pp;
    memset(&p, 0, sizeof(p));
    p.g_first = g_first; p.g_field = g_field; p.g_flat = g_flat; p.g_fm = g_fm; p.fe = field_emb; p.fm_sum = fm_sum;
    p.aux = aux; p.A = plan->A;
    p.pos = reinterpret_cast<const long long*>(positions); p.B = batch; p.S = plan->S; p.T = plan->T;
    p.F = plan->n_fields; p.D = plan->fm_dim; p.dmax = plan->max_tdim;
    p.send_slots = send_slots; p.n_sent = n_sent; p.grad_scale = grad_scale;
    if (send_slots && n_sent <= 0) return DFM_OK;
    fill_slot_tf(plan, p.slot_tf);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool v4 = plan->vec == 4 && al16(g_flat) && al16(g_field) && al16(fm_sum) && al16(g_vec) && al16(field_emb);
    const int lanes = p.dmax / (v4 ? 4 : 1);
    DFM_REQUIRE(lanes <= 32, DFM_ERR_UNSUPPORTED, "dfm_shard_pack_grad: table dim %d too wide", p.dmax);
    const int G = next_pow2(lanes);
    long long blocks = ceil_div(send_slots ? n_sent : batch * plan->S, 256 / G);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (v4) shard_pack_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*a, p, G, g_vec, *pd);
    else shard_pack_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*a, p, G, g_vec, *pd);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_shard_pack_grad(const dfm_plan* plan, int64_t batch, const int64_t* positions, const float* g_first,
                        const float* g_field, const float* g_flat, const float* g_fm, const float* fm_sum,
                        const float* field_emb, const uint32_t* aux, float* g_vec, float grad_scale, void* stream) {
    return shard_pack_impl(plan, batch, positions, g_first, g_field, g_flat, g_fm, fm_sum, field_emb, aux, g_vec, 0, nullptr,
                           nullptr, nullptr, 0, grad_scale, stream);
}

int dfm_shard_pack_grad_p2p(const dfm_plan* plan, int64_t batch, const int64_t* positions, const float* g_first,
                            const float* g_field, const float* g_flat, const float* g_fm, const float* fm_sum,
                            const float* field_emb, const uint32_t* aux, int n_peers, const int64_t* peer_start,
                            float* const* peer_rows, const uint32_t* send_slots, float grad_scale, void* stream) {
    DFM_REQUIRE(n_peers > 0, DFM_ERR_INVALID, "dfm_shard_pack_grad_p2p: no peers");
    return shard_pack_impl(plan, batch, positions, g_first, g_field, g_flat, g_fm, fm_sum, field_emb, aux, nullptr, n_peers,
                           peer_start, peer_rows, send_slots, send_slots ? peer_start[n_peers] - peer_start[0] : 0, grad_scale,
                           stream);
}

size_t dfm_shard_route_workspace_bytes(const dfm_plan* plan, int64_t batch) {
    if (!plan || batch < 0) return 0;
    const long long nblk = ceil_div((long long)batch * (plan->S > 0 ? plan->S : 1), RT_TILE) + 1;
    return (size_t)nblk * RT_MAXW * 4 + (size_t)nblk * RT_MAXW * 8 + 512;
}

int dfm_shard_route(const dfm_plan* plan, int world, const int64_t* global_row_base, int64_t batch,
                    const void* const* inputs, uint32_t* send_keys, int64_t* positions, int64_t* counts,
                    uint32_t* send_slots, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(plan && global_row_base && inputs && counts && world > 0 && world <= RT_MAXW && batch >= 0, DFM_ERR_INVALID,
                "dfm_shard_route: bad argument (world must be 1..%d)", RT_MAXW);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n = (long long)batch * plan->S;
    if (n == 0) { DFM_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)world * 8, st)); return DFM_OK; }
    DFM_REQUIRE(send_keys && positions && workspace, DFM_ERR_INVALID, "dfm_shard_route: null tensor");
    DFM_REQUIRE(workspace_bytes >= dfm_shard_route_workspace_bytes(plan, batch), DFM_ERR_WORKSPACE, "dfm_shard_route: workspace too small");
    RouteArgs* a = new RouteArgs;
    struct Gd { RouteArgs* p; ~Gd() { delete p; } } gd{a};
    memset(a, 0, sizeof(*a));
    a->S = plan->S; a->world = world; a->B = batch; a->status = status;
    int nt = 0;
    for (int f = 0; f < plan->n_fields; ++f) {
        if (!shard_field(plan, f, false)) continue;
        DFM_REQUIRE(inputs[f], DFM_ERR_INVALID, "dfm_shard_route: field %d has no id column", f);
        RouteField& rf = a->f[nt++];
        rf.ids = static_cast<const long long*>(inputs[f]);
        rf.gbase = (unsigned)global_row_base[f];
        rf.slot_base = plan->slot_base[f]; rf.max_len = plan->max_len[f];
        rf.bag = plan->kind[f] == DFM_SEQUENCE ? 1 : 0;
        rf.rot = f % world;
        rf.vocab = global_row_base[f + 1] - global_row_base[f];
    }
    fill_slot_tf(plan, a->slot_tf);
    const int nblk = (int)ceil_div(n, RT_TILE);
    int* block_counts = static_cast<int*>(workspace);
    long long* offsets = reinterpret_cast<long long*>(static_cast<char*>(workspace) + align_up((size_t)nblk * RT_MAXW * 4, 256));
    route_count_kernel<<<nblk, 256, 0, st>>>(*a, block_counts);
    route_scan_kernel<<<1, 32, 0, st>>>(block_counts, nblk, world, offsets, reinterpret_cast<long long*>(counts));
    route_scatter_kernel<<<nblk, 256, 0, st>>>(*a, offsets, nblk, send_keys, reinterpret_cast<long long*>(positions), send_slots);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

size_t dfm_shard_unique_workspace_bytes(int64_t n) {
    const long long nblk = ceil_div(n > 0 ? n : 1, UQ_TILE) + 1;
    return (size_t)nblk * 4 + (size_t)nblk * RT_MAXW * 4 + (size_t)nblk * 8 + 1024;
}

int dfm_shard_ukeys(const dfm_plan* plan, int world, const int64_t* vbase, const int64_t* vocab, int lbits, int64_t batch,
                    const void* const* inputs, uint32_t* keys, uint32_t* payload, int64_t* positions, int32_t* status,
                    void* stream) {
    DFM_REQUIRE(plan && vbase && vocab && inputs && world > 0 && world <= RT_MAXW && batch >= 0 && lbits > 0 && lbits < 31,
                DFM_ERR_INVALID, "dfm_shard_ukeys: bad argument (world must be 1..%d)", RT_MAXW);
    if (batch == 0) return DFM_OK;
    DFM_REQUIRE(keys && payload && positions, DFM_ERR_INVALID, "dfm_shard_ukeys: null tensor");
    UKeyArgs* a = new UKeyArgs;
    struct Gd { UKeyArgs* p; ~Gd() { delete p; } } gd{a};
    memset(a, 0, sizeof(*a));
    int nt = 0, ts = 0, pbits = 0;
    while ((1 << pbits) < plan->S) ++pbits;
    for (int f = 0; f < plan->n_fields; ++f) {
        if (!shard_field(plan, f, false)) continue;
        DFM_REQUIRE(inputs[f], DFM_ERR_INVALID, "dfm_shard_ukeys: field %d has no id column", f);
        UKeyField& kf = a->f[nt];
        kf.ids = static_cast<const long long*>(inputs[f]);
        kf.vbase = (unsigned)vbase[f];
        kf.slot_base = plan->slot_base[f]; kf.max_len = plan->max_len[f];
        kf.bag = plan->kind[f] == DFM_SEQUENCE ? 1 : 0;
        kf.rot = f % world;
        kf.vocab = vocab[f];
        for (int l = 0; l < plan->max_len[f]; ++l) { a->ts_field[ts] = (unsigned short)nt; a->ts_pos[ts] = (unsigned short)l; ++ts; }
        ++nt;
    }
    DFM_REQUIRE(ts > 0, DFM_ERR_INVALID, "dfm_shard_ukeys: the plan has no sharded (foreign) table");
    DFM_REQUIRE(((long long)batch << pbits) < 0xffffffffLL, DFM_ERR_UNSUPPORTED, "dfm_shard_ukeys: batch * slots must fit 32 bits");
    a->S_sh = ts; a->world = world; a->lbits = lbits; a->pbits = pbits; a->B = batch; a->status = status;
    a->pad = (unsigned)world << lbits;
    const long long n = (long long)batch * ts;
    long long blocks = ceil_div(n, 256);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    shard_ukeys_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a, keys, payload, reinterpret_cast<long long*>(positions));
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_shard_unique(const dfm_plan* plan, int world, int lbits, int64_t batch, int64_t n, const uint32_t* sorted_keys,
                     const uint32_t* sorted_payload, uint32_t* unique_keys, uint32_t* unique_index, int64_t* positions,
                     int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(plan && counts && world > 0 && world <= RT_MAXW && n >= 0, DFM_ERR_INVALID, "dfm_shard_unique: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) { DFM_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)world * 8, st)); return DFM_OK; }
    DFM_REQUIRE(sorted_keys && sorted_payload && unique_keys && unique_index && positions && workspace, DFM_ERR_INVALID, "dfm_shard_unique: null tensor");
    DFM_REQUIRE(workspace_bytes >= dfm_shard_unique_workspace_bytes(n), DFM_ERR_WORKSPACE, "dfm_shard_unique: workspace too small");
    UniqArgs* a = new UniqArgs;
    struct Gd { UniqArgs* p; ~Gd() { delete p; } } gd{a};
    memset(a, 0, sizeof(*a));
    for (int s = 0; s < plan->S; ++s) {
        const int f = plan->slot_field[s];
        a->slot_sb[s] = (unsigned short)plan->slot_base[f];
        a->slot_ml[s] = (unsigned short)plan->max_len[f];
    }
    int pbits = 0;
    while ((1 << pbits) < plan->S) ++pbits;
    a->B = batch; a->pbits = pbits; a->pad = (unsigned)world << lbits; a->lmask = (1u << lbits) - 1u;
    const int nblk = (int)ceil_div(n, UQ_TILE);
    char* ws = static_cast<char*>(workspace);
    int* blk_heads = reinterpret_cast<int*>(ws);
    int* blk_owner = reinterpret_cast<int*>(ws + align_up((size_t)nblk * 4, 256));
    long long* blk_base = reinterpret_cast<long long*>(ws + align_up((size_t)nblk * 4, 256) + align_up((size_t)nblk * RT_MAXW * 4, 256));
    uniq_count_kernel<<<nblk, 256, 0, st>>>(sorted_keys, n, a->pad, lbits, blk_heads, blk_owner);
    uniq_scan_kernel<<<1, 32, 0, st>>>(blk_heads, blk_owner, nblk, world, blk_base, reinterpret_cast<long long*>(counts));
    uniq_emit_kernel<<<nblk, 256, 0, st>>>(*a, sorted_keys, sorted_payload, n, blk_base, unique_keys, unique_index,
                                           reinterpret_cast<long long*>(positions));
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_shard_gather2(const dfm_plan* local_plan, int world, int rank, int64_t n_keys, const uint32_t* local_keys,
                      const float* const* params, int n_peers, const int64_t* peer_start, float* const* peer_vec,
                      float* const* peer_sc, uint32_t* backward_keys, void* stream) {
    DFM_REQUIRE(local_plan && params && world > 0 && rank >= 0 && rank < world, DFM_ERR_INVALID, "dfm_shard_gather2: bad argument");
    if (n_keys <= 0) return DFM_OK;
    DFM_REQUIRE(local_keys && backward_keys && n_peers >= 1 && n_peers <= RT_MAXW_ && peer_start && peer_vec && peer_sc, DFM_ERR_INVALID,
                "dfm_shard_gather2: null tensor / 1..%d peers", RT_MAXW_);
    DFM_REQUIRE(peer_start[n_peers] - peer_start[0] == n_keys, DFM_ERR_INVALID, "dfm_shard_gather2: the peer segments must cover the %lld keys", (long long)n_keys);
    ShardArgs* a = new ShardArgs;
    PeerDst2* pd = new PeerDst2;
    struct Gd { ShardArgs* p; PeerDst2* q; ~Gd() { delete p; delete q; } } gd{a, pd};
    std::vector<int64_t> zeros(local_plan->n_fields + 1, 0);
    int rc = fill_shard_args(local_plan, zeros.data(), params, world, rank, *a, true);
    if (rc) return rc;
    memset(pd, 0, sizeof(*pd));
    for (int q = 0; q < n_peers; ++q) {
        DFM_REQUIRE(peer_vec[q] && peer_sc[q] && ((reinterpret_cast<uintptr_t>(peer_vec[q]) | reinterpret_cast<uintptr_t>(peer_sc[q])) & 15u) == 0 &&
                    peer_start[q] <= peer_start[q + 1], DFM_ERR_INVALID, "dfm_shard_gather2: peer %d has a null / unaligned buffer or a negative segment", q);
        pd->start[q] = peer_start[q]; pd->vec[q] = peer_vec[q]; pd->sc[q] = peer_sc[q];
    }
    pd->start[n_peers] = peer_start[n_peers];
    pd->n = n_peers;
    const bool v4 = local_plan->vec == 4;
    const int lanes = a->dmax / (v4 ? 4 : 1);
    DFM_REQUIRE(lanes <= 32, DFM_ERR_UNSUPPORTED, "dfm_shard_gather2: table dim %d too wide", a->dmax);
    const int G = next_pow2(lanes);
    long long blocks = ceil_div(n_keys, 256 / G);
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (v4) shard_gather2_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(*a, n_keys, local_keys, G, backward_keys, *pd);
    else shard_gather2_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(*a, n_keys, local_keys, G, backward_keys, *pd);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_shard_push_counts(const int64_t* counts, int world, int rank, int64_t* const* peer_matrix, void* stream) {
    DFM_REQUIRE(counts && peer_matrix && world >= 1 && world <= RT_MAXW_ && rank >= 0 && rank < world, DFM_ERR_INVALID,
                "dfm_shard_push_counts: 1..%d ranks", RT_MAXW_);
    PeerI64 d;
    memset(&d, 0, sizeof(d));
    for (int r = 0; r < world; ++r) {
        DFM_REQUIRE(peer_matrix[r], DFM_ERR_INVALID, "dfm_shard_push_counts: null matrix of rank %d", r);
        d.p[r] = reinterpret_cast<long long*>(peer_matrix[r]);
    }
    push_counts_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const long long*>(counts), world, rank, d);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_shard_push_keys(const uint32_t* send_keys, const int64_t* matrix, int world, int rank, int64_t capacity,
                        uint32_t* const* peer_keys, void* stream) {
    DFM_REQUIRE(send_keys && matrix && peer_keys && world >= 1 && world <= RT_MAXW_ && rank >= 0 && rank < world && capacity > 0,
                DFM_ERR_INVALID, "dfm_shard_push_keys: bad argument (1..%d ranks)", RT_MAXW_);
    PeerU32 d;
    memset(&d, 0, sizeof(d));
    for (int r = 0; r < world; ++r) {
        DFM_REQUIRE(peer_keys[r], DFM_ERR_INVALID, "dfm_shard_push_keys: null key buffer of rank %d", r);
        d.p[r] = peer_keys[r];
    }
    push_keys_kernel<<<2 * sm_count(), 256, 0, static_cast<cudaStream_t>(stream)>>>(send_keys, reinterpret_cast<const long long*>(matrix),
                                                                                   world, rank, capacity, d);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

int dfm_peer_barrier(uint32_t* const* peer_flags, int world, int rank, uint32_t epoch, void* stream) {
    DFM_REQUIRE(peer_flags && world >= 1 && world <= RT_MAXW_ && rank >= 0 && rank < world, DFM_ERR_INVALID,
                "dfm_peer_barrier: 1..%d ranks", RT_MAXW_);
    PeerFlags pf;
    memset(&pf, 0, sizeof(pf));
    for (int r = 0; r < world; ++r) {
        DFM_REQUIRE(peer_flags[r], DFM_ERR_INVALID, "dfm_peer_barrier: null flag array of rank %d", r);
        pf.flags[r] = peer_flags[r];
    }
    peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pf, world, rank, epoch);
    DFM_CHECK_LAUNCH();
    return DFM_OK;
}

}  // extern "C"
