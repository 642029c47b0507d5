/*
 * deepfm_b200.h -- C ABI of libdeepfm_b200.so: the sm_100a CUDA kernels behind the drop-in
 * FeatureEmbedding / FMInteraction / CIN / MultiHeadSelfAttention modules.
 *
 * The reference (CodexploreRepo/deepfm) has no FFI: its hot path is PyTorch nn.Modules calling
 * ATen.  The boundary a maintainer binds instead is this header -- one entry point per ATen
 * call-site cluster the kernels replace (cited per function as reference file:line).
 * INTEGRATION.md shows the ctypes stub and the autograd.Function that sit on top.
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, an opaque host-side plan; no torch types.
 *   - every function returns DFM_OK (0) or a negative dfm_status; dfm_last_error() gives the
 *     message for the calling thread.  Nothing throws across the boundary.
 *   - the CALLER owns every buffer (outputs, workspaces): nothing is allocated or freed here
 *     except the host-side plan objects.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no internal sync.
 *   - fp32 values, int64 ids, row-major contiguous tensors, exactly as the reference modules
 *     produce/consume them (deepfm/data/dataset.py:32-37).
 */
#ifndef DEEPFM_B200_H
#define DEEPFM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DFM_API __attribute__((visibility("default")))
#else
#define DFM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dfm_status {
    DFM_OK = 0,
    DFM_ERR_INVALID = -1,      /* bad argument (null pointer, negative size, unknown enum) */
    DFM_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels are instantiated for */
    DFM_ERR_CUDA = -3,         /* a CUDA runtime call failed; see dfm_last_error() */
    DFM_ERR_WORKSPACE = -4     /* workspace smaller than dfm_*_workspace_bytes() */
} dfm_status;

/* Field kinds / combiners: deepfm/data/schema.py:7-21 (FeatureType, FieldSchema.combiner). */
enum { DFM_SPARSE = 0, DFM_SEQUENCE = 1, DFM_DENSE = 2 };
enum { DFM_SUM = 0, DFM_MEAN = 1, DFM_MAX = 2 };
/* Gradient layout of the embedding tables produced by dfm_embed_bwd. */
enum { DFM_GRAD_DENSE = 0, DFM_GRAD_ROWSPARSE = 1, DFM_GRAD_SKIP_TABLES = 2 /* only DENSE-field / projection grads */,
       /* flag, OR-ed into the mode of dfm_embed_bwd / dfm_rows_bwd: sorted_keys / sorted_payload already hold the
        * result of dfm_sort_keys for this batch (the device-side input pipeline sorts ahead of the step) */
       DFM_GRAD_PRESORTED = 0x100 };

DFM_API const char* dfm_last_error(void);
DFM_API int dfm_version(void);
/* Device properties the host side sizes grids with: out[0]=SM count, out[1]=cc major,
 * out[2]=cc minor, out[3]=max opt-in shared memory per block. */
DFM_API int dfm_device_info(int device, int64_t out[4]);

/* ------------------------------------------------------------------------------------------
 * Embedding plan: the immutable per-module description of FeatureEmbedding.__init__
 * (deepfm/models/layers/embedding.py:20-64).  Host memory only.
 * ---------------------------------------------------------------------------------------- */
typedef struct dfm_plan dfm_plan;

DFM_API dfm_plan* dfm_plan_create(int n_fields, const int32_t* kind, const int32_t* dim,
                          const int64_t* vocab, const int32_t* max_len,
                          const int32_t* combiner, int fm_dim);
DFM_API void dfm_plan_destroy(dfm_plan* plan);

/* Derived sizes.  out[0]=T (total_embedding_dim), out[1]=S (id slots per sample: 1 per SPARSE
 * field, max_len per SEQUENCE field), out[2]=total table rows (sum of vocab over id fields;
 * also the PAD key), out[3]=aux 4-byte words per sample (bag scales / arg-max positions),
 * out[4]=1 if flat and field_embeddings are byte-identical (all dims == fm_dim, no projection),
 * out[5]=max table dim, out[6]=number of sort key bits, out[7]=vector width (4 or 1). */
DFM_API int dfm_plan_info(const dfm_plan* plan, int64_t out[8]);

/* Integer artefacts (bit-exact contracts, oracle: slot_layout / emit_keys). */
DFM_API int dfm_plan_slots(const dfm_plan* plan, int32_t* slot_field, int32_t* slot_pos,
                   int64_t* row_base /* n_fields+1 */);

/* ------------------------------------------------------------------------------------------
 * K1  FeatureEmbedding.forward fused with FMInteraction.forward
 *     (embedding.py:76-126 and fm.py:18-23; ATen: embedding, embedding_bag, linear, stack,
 *      sum, cat, pow, sub, mul).
 *
 *   inputs[f] : device pointer to the field's batch column -- int64 (B,) SPARSE,
 *               int64 (B, max_len) zero-padded SEQUENCE, float (B,) DENSE.
 *   params[5f+0..4] : second-order weight, second-order bias (DENSE only), first-order weight,
 *               first-order bias (DENSE only), projection weight (D, d_f) or NULL.
 *   first_order (B,1), field_emb (B,F,D), flat (B,T): outputs.  field_emb may equal flat when
 *               dfm_plan_info()[4] == 1; the tensor is then written once.
 *   fm_out (B,1), fm_sum (B,D): optional (NULL to skip) FM value and per-dim field sum
 *               (kept for the backward).
 *   keys (B*S) uint32: optional; global row of every id slot, PAD key for id 0 (input of the
 *               sorted backward).   aux (B*A) 4-byte words: required iff A > 0.
 *   status: optional device int32, set to 1 if an id was outside [0, vocab) (the id is then
 *               treated as padding; the reference raises IndexError on CPU).
 * ---------------------------------------------------------------------------------------- */
DFM_API int dfm_embed_fwd(const dfm_plan* plan, int64_t batch, const void* const* inputs,
                  const float* const* params, float* first_order, float* field_emb,
                  float* flat, float* fm_out, float* fm_sum, uint32_t* keys, uint32_t* aux,
                  int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  backward of K1 w.r.t. every FeatureEmbedding parameter, fused with the FM backward
 *     (g_e = g_field + g_fm * (fm_sum - e)) and with the L2 penalty gradient 2*l2*p of
 *     BaseCTRModel.get_l2_reg_loss (base.py:78-83).  Replaces ATen embedding_dense_backward,
 *     _embedding_bag_dense_backward, linear backward.
 *
 *   Sorted-index segmented reduction: keys are radix-sorted (stable, payload = b*S+slot), equal
 *   keys form a segment, each segment is summed in payload order by exactly one thread group;
 *   segments longer than one chunk are stitched by a second fixed-order pass.  No float
 *   atomics; bit-reproducible.
 *
 *   g_first (B,1), g_field (B,F,D), g_flat (B,T), g_fm (B,1): upstream gradients, any may be NULL.
 *   l2 : lambda (0 disables).  l2_gscale: optional device scalar multiplying the L2 gradient
 *        (the upstream gradient of the penalty term; NULL = 1).
 *   mode DFM_GRAD_DENSE: grads[5f+k] are dense gradients with the shapes of params[5f+k]
 *        (reference semantics: every element receives 2*l2*p, row 0 gets no lookup gradient).
 *   mode DFM_GRAD_ROWSPARSE: table entries of grads[] are ignored; instead
 *        sorted_keys/sorted_payload (B*S) hold the sorted pairs and, at the first position of
 *        every segment, row_grad2 (B*S, max table dim) / row_grad1 (B*S) hold the row's summed
 *        gradient plus 2*l2*w[row] (L2 on touched rows only).  n_valid (device int64[2]) gets
 *        {number of non-PAD keys, number of unique rows}.  Non-table entries of grads[] (DENSE
 *        field Linears, projections) are always written densely.
 * ---------------------------------------------------------------------------------------- */
DFM_API size_t dfm_embed_bwd_workspace_bytes(const dfm_plan* plan, int64_t batch);

DFM_API int dfm_embed_bwd(const dfm_plan* plan, int64_t batch, const void* const* inputs,
                  const float* const* params, const float* g_first, const float* g_field,
                  const float* g_flat, const float* g_fm, const float* field_emb,
                  const float* flat, const float* fm_sum, const uint32_t* keys,
                  const uint32_t* aux, float l2, const float* l2_gscale, int mode,
                  float* const* grads, uint32_t* sorted_keys, uint32_t* sorted_payload,
                  float* row_grad2, float* row_grad1, int64_t* n_valid, void* workspace,
                  size_t workspace_bytes, void* stream);

/* Stand-alone pieces of K2, exposed so the integer artefacts can be checked bit-exactly
 * against the oracle (emit_keys / sort_pairs / segment_heads). */
DFM_API int dfm_emit_keys(const dfm_plan* plan, int64_t batch, const void* const* inputs,
                  uint32_t* keys, void* stream);
DFM_API size_t dfm_sort_keys_workspace_bytes(const dfm_plan* plan, int64_t n);
DFM_API int dfm_sort_keys(const dfm_plan* plan, int64_t n, const uint32_t* keys, uint32_t* sorted_keys,
                  uint32_t* sorted_payload, void* workspace, size_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * FMInteraction stand-alone (fm.py:18-23) for inputs that do not come out of K1.
 * ---------------------------------------------------------------------------------------- */
DFM_API int dfm_fm_fwd(const float* e, int64_t batch, int n_fields, int dim, float* out, void* stream);
DFM_API int dfm_fm_bwd(const float* e, const float* g_out, int64_t batch, int n_fields, int dim,
               float* g_e, void* stream);

/* ------------------------------------------------------------------------------------------
 * lambda * sum_p ||p||^2 over a list of tensors (base.py:78-83; ATen norm + pow per param).
 * Deterministic two-stage reduction; `out` is one device float; workspace >= 4096 floats.
 * ---------------------------------------------------------------------------------------- */
DFM_API int dfm_sumsq(int n_tensors, const float* const* ptrs, const int64_t* numel, float scale,
              float* out, float* workspace, void* stream);
/* Incrementally maintained ||W||^2 (SURVEY hard part 2: the reference re-reads every table every step):
 * dfm_sumsq_acc writes the exact sum of squares of the tensor list into the device double acc[0] (the audited slow
 * path, same reduction as dfm_sumsq); dfm_adam_rows then adds sum(w_new^2 - w_old^2) of the rows it touches, so the
 * value stays current without another pass over the tables; dfm_l2_combine forms out[0] = lam * (a[0] + b[0])
 * (b may be NULL) -- the scalar get_l2_reg_loss returns. */
DFM_API int dfm_sumsq_acc(int n_tensors, const float* const* ptrs, const int64_t* numel, double* acc,
                          float* workspace, void* stream);
DFM_API int dfm_l2_combine(const double* acc_a, const double* acc_b, float lam, float* out, void* stream);
/* g[i] (+)= coef * scale_dev[0] * p[i]  (the dense L2 gradient when K2 is not in the graph). */
DFM_API int dfm_axpy(const float* p, int64_t numel, float coef, const float* scale_dev, float* g,
             int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * CIN (deepfm/models/layers/cin.py:26-105; ATen einsum + reshape + conv1d(k=1) + relu + split +
 * sum).  layer_sizes[i] = L_i; weights[i] (L_i, K_i) with K_i = H_i * F, channel k = h*F + f;
 * split is [direct first, next second] (cin.py:93-96).  The (B, H*F, D) outer product is never
 * written to memory.  precision 0 = fp32 CUDA cores (reference-exact up to summation order); precision 1 = TF32 on the
 * tcgen05 tensor cores, forward AND backward (fp32 accumulation in tensor memory; 2e-3 forward / 3e-3 gradient tolerance),
 * for layers the tensor-core tiles cover (F <= 64, L <= 256, D % 4 == 0), CUDA cores otherwise.
 *   dfm_cin_sizes: out[0] = output_dim, out[1] = bytes of `acts` (post-ReLU activations of every
 *   layer, kept for the backward), out[2] = bytes of the backward workspace, out[3] = activation
 *   floats per sample.
 * ---------------------------------------------------------------------------------------- */
DFM_API int dfm_cin_sizes(int n_fields, int dim, int n_layers, const int32_t* layer_sizes,
                          int split_half, int64_t batch, int64_t out[4]);
DFM_API int dfm_cin_fwd(const float* x0, int64_t batch, int n_fields, int dim, int n_layers,
                        const int32_t* layer_sizes, int split_half, const float* const* weights,
                        const float* const* biases, int precision, float* out, float* acts,
                        void* stream);
DFM_API int dfm_cin_bwd(const float* x0, const float* g_out, int64_t batch, int n_fields, int dim,
                        int n_layers, const int32_t* layer_sizes, int split_half,
                        const float* const* weights, int precision, const float* acts, float* g_x0,
                        float* const* g_weights, float* const* g_biases, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * One _AttentionBlock (deepfm/models/layers/attention.py:70-120) fused in shared memory.
 * params[0..9] = W_q.weight (A,D), W_q.bias, W_k.weight, W_k.bias, W_v.weight, W_v.bias,
 * W_out.weight (D,A), W_out.bias, layer_norm.weight, layer_norm.bias (last two only when
 * use_residual).  The backward recomputes the forward from x; g_params has the same order.
 * ---------------------------------------------------------------------------------------- */
DFM_API size_t dfm_attn_workspace_bytes(int64_t batch, int n_fields, int dim, int attention_dim,
                                        int heads);
DFM_API int dfm_attn_fwd(const float* x, int64_t batch, int n_fields, int dim, int attention_dim,
                         int heads, int use_residual, const float* const* params, float* out,
                         void* stream);
DFM_API int dfm_attn_bwd(const float* x, const float* g_out, int64_t batch, int n_fields, int dim,
                         int attention_dim, int heads, int use_residual,
                         const float* const* params, float* g_x, float* const* g_params,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-sharded tables over W GPUs (new; the reference is single-device).  owner(f, id) = (id + f) mod W
 * (f = schema index of the field: spreads the hot small ids of the tables over the ranks),
 * local_row = id div W, per field (oracle: shard_route).  The NCCL all-to-all of keys / vectors /
 * vector gradients is issued by the host side between these calls.
 *   Exchanged rows are d_max + 4 floats wide (16-byte aligned): one collective carries the vector and
 *   its scalars.
 *   dfm_shard_route : sample side.  Stable grouping of the B*S id slots (source order b*S + s) by
 *       owner: send_keys (<= B*S) global rows in send order, counts (W), and positions: for field f
 *       the (B, max_len[f]) block at offset B * slot_base[f] holds the 1-based send position of every
 *       slot (0 = nothing sent: a padding id of a multi-hot bag; the padding id of a SPARSE field IS
 *       sent, its row 0 is returned as stored).  The blocks are K1's id columns of the sample-side plan.
 *       send_slots (optional, <= B*S): the inverse map, send position -> id slot index b*S + s; with it
 *       dfm_shard_pack_grad_p2p walks the send order, so every peer receives one sequential store stream.
 *       An id outside [0, vocabulary_size) (global_row_base[f+1] - global_row_base[f]) is routed as the padding
 *       id 0 and sets *status = 1 (optional device word; the reference raises IndexError).
 *   dfm_shard_gather : owner side.  keys = global rows (row_base[f] + id) as received; writes the
 *       reply rows [row (d), first-order weight, 0, 0, 0] and the local sort keys
 *       (local_row_base[f] + local_row, PAD for id 0) the owner-side backward consumes.
 *   dfm_plan_set_table_stride : lets K1 (dfm_embed_fwd) read every plain SPARSE / sum- or mean-bag
 *       table with a row stride != dim, i.e. straight out of the received reply rows (row 0 of that
 *       buffer is a reserved zero row, reply i is row i + 1; first-order at column d).
 *   dfm_shard_pack_grad : sample side.  positions as above; writes the gradient rows in send order:
 *       SPARSE [g_flat + g_field + g_fm * fm_sum (d), g_first, g_fm, 0, 0];  bag member
 *       [scale * (g_flat + g_field + g_fm * (fm_sum - e_bag)) (d), scale * g_first, 0, 0, 0] with
 *       scale = 1 (sum) or 1 / non-pad count (mean, from the forward's aux record); field_emb / aux
 *       may be null when the plan has no bag fields.  grad_scale multiplies every packed value: 1 / world makes
 *       the owner's summed table gradient the gradient of the GLOBAL-mean loss, the scale the averaged
 *       data-parallel parameters have (the L2 term 2 l2 w is added once, by the owner, unscaled).
 *   dfm_rows_bwd : owner side backward = K2 on a row list: keys (n) with one packed gradient row
 *       each; same sort / segreduce / stitch kernels and modes as dfm_embed_bwd, the field is
 *       derived from the key and -(sum g_fm) w[row] + 2 l2 w[row] is folded in at the segment end.
 * ---------------------------------------------------------------------------------------- */
DFM_API size_t dfm_shard_route_workspace_bytes(const dfm_plan* plan, int64_t batch);
DFM_API int dfm_shard_route(const dfm_plan* plan, int world, const int64_t* global_row_base,
                            int64_t batch, const void* const* inputs, uint32_t* send_keys,
                            int64_t* positions, int64_t* counts, uint32_t* send_slots, int32_t* status,
                            void* workspace, size_t workspace_bytes, void* stream);
DFM_API int dfm_shard_gather(const dfm_plan* local_plan, int world, int rank,
                             const int64_t* global_row_base, int64_t n_keys, const uint32_t* keys,
                             const float* const* params, float* rows, uint32_t* local_keys,
                             void* stream);
DFM_API int dfm_shard_pack_grad(const dfm_plan* plan, int64_t batch, const int64_t* positions,
                                const float* g_first, const float* g_field, const float* g_flat,
                                const float* g_fm, const float* fm_sum, const float* field_emb,
                                const uint32_t* aux, float* g_rows, float grad_scale, void* stream);
/* Fused exchange over peer memory (NVLink P2P stores; peer buffers come from CUDA IPC / torch symmetric
 * memory): the same two kernels write every row straight into the destination GPU's buffer instead of a
 * local staging buffer followed by an all-to-all.  Rows [peer_start[p], peer_start[p+1]) of this rank's send
 * order (dfm_shard_gather_p2p: of the received keys, which are grouped by source) go to
 * peer_rows[p] + (i - peer_start[p]) * (d_max + 4).  The caller separates the writes from the consumer with
 * a cross-GPU barrier. */
DFM_API int dfm_shard_gather_p2p(const dfm_plan* local_plan, int world, int rank,
                                 const int64_t* global_row_base, int64_t n_keys, const uint32_t* keys,
                                 const float* const* params, int n_peers, const int64_t* peer_start,
                                 float* const* peer_rows, uint32_t* local_keys, void* stream);
DFM_API int dfm_shard_pack_grad_p2p(const dfm_plan* plan, int64_t batch, const int64_t* positions,
                                    const float* g_first, const float* g_field, const float* g_flat,
                                    const float* g_fm, const float* fm_sum, const float* field_emb,
                                    const uint32_t* aux, int n_peers, const int64_t* peer_start,
                                    float* const* peer_rows, const uint32_t* send_slots, float grad_scale,
                                    void* stream);
/* Exchange of UNIQUE rows (v2; what ShardedFeatureEmbedding uses).  Under CTR-like skew most id slots of a batch
 * repeat a row the same GPU already asked the same owner for, so every distinct (owner, row) travels once per step
 * and direction.  Local tables are sized ceil(V / W) rows on every rank, so the owner's sort key of a row,
 * lkey = vbase[f] + id div W (vbase = exclusive prefix sum of ceil(V / W) over the sharded fields), is the same number
 * on every rank.
 *   dfm_shard_ukeys  : sample side.  Per sharded id slot (compact index j = b * S_sh + ts): key = owner << lbits | lkey
 *       (PAD = world << lbits for an unsent bag padding entry), payload = b << pbits | plan slot (the layout K2 decodes);
 *       positions (the plan's (B, max_len) blocks, see dfm_shard_route) are zeroed; out-of-range ids -> id 0 + *status.
 *   dfm_sort_pairs   : stable radix sort of (key, payload) on the low `bits` bits (CUB).
 *   dfm_shard_unique : sorted keys -> unique_keys (lkeys in send order: grouped by owner, ascending), counts (W),
 *       unique_index[p] = ordinal of position p's key, positions[slot] = 1 + that ordinal (row of the reply buffer).
 *   dfm_shard_gather2: owner side.  local_keys (lkeys as received, grouped by source) -> vector rows (d floats) and
 *       scalar rows [first-order weight, 0, 0, 0] stored into the requesters' buffers: row i of source segment
 *       [peer_start[p], peer_start[p+1]) goes to peer_vec[p] + (i - start) * d / peer_sc[p] + (i - start) * 4 (peer
 *       memory, or a local staging buffer with n_peers = 1); backward_keys = lkey, PAD for the padding row id 0.
 *   dfm_shard_bwd_peer: sample side backward.  K2's segmented reduction over the sample-side sorted (key, payload)
 *       stream: the gradient rows of all slots that share a key are summed in sorted order and ONE row per unique
 *       key -- [sum scale * (g_flat + g_field + g_fm * (fm_sum [- e_bag]))] and [sum g_first, sum g_fm, 0, 0], times
 *       grad_scale -- is stored into the owner's buffers at the key's position in the send order.
 *   dfm_rows_bwd with g_scalars != NULL consumes that split layout on the owner ((n, d) vectors + (n, 4) scalars). */
DFM_API int dfm_shard_ukeys(const dfm_plan* plan, int world, const int64_t* vbase, const int64_t* vocab, int lbits,
                            int64_t batch, const void* const* inputs, uint32_t* keys, uint32_t* payload,
                            int64_t* positions, int32_t* status, void* stream);
DFM_API size_t dfm_sort_pairs_workspace_bytes(int64_t n, int bits);
DFM_API int dfm_sort_pairs(int64_t n, int bits, const uint32_t* keys, const uint32_t* payload, uint32_t* sorted_keys,
                           uint32_t* sorted_payload, void* workspace, size_t workspace_bytes, void* stream);
DFM_API size_t dfm_shard_unique_workspace_bytes(int64_t n);
DFM_API int dfm_shard_unique(const dfm_plan* plan, int world, int lbits, int64_t batch, int64_t n,
                             const uint32_t* sorted_keys, const uint32_t* sorted_payload, uint32_t* unique_keys,
                             uint32_t* unique_index, int64_t* positions, int64_t* counts, void* workspace,
                             size_t workspace_bytes, void* stream);
DFM_API int dfm_shard_gather2(const dfm_plan* local_plan, int world, int rank, int64_t n_keys,
                              const uint32_t* local_keys, const float* const* params, int n_peers,
                              const int64_t* peer_start, float* const* peer_vec, float* const* peer_sc,
                              uint32_t* backward_keys, void* stream);
DFM_API int dfm_shard_bwd_peer(const dfm_plan* plan, int64_t batch, const float* g_first, const float* g_field,
                               const float* g_flat, const float* g_fm, const float* field_emb, const float* fm_sum,
                               const uint32_t* aux, const uint32_t* sorted_keys, const uint32_t* sorted_payload,
                               const uint32_t* unique_index, int64_t n_sorted, uint32_t pad_key, int n_peers,
                               const int64_t* peer_start, float* const* peer_vec, float* const* peer_sc,
                               float grad_scale, void* workspace, size_t workspace_bytes, void* stream);
/* Device-side barrier between the ranks' streams over peer-mapped memory (no reference counterpart: the reference is
 * single-device).  peer_flags[r] = device address, valid on THIS GPU, of rank r's flag array (>= world uint32 words,
 * zero-initialised, in memory every rank has mapped: the symmetric exchange allocation).  Enqueues one tiny kernel on
 * `stream`: it publishes `epoch` to every rank and waits until every rank has published an epoch >= `epoch`; all peer
 * stores enqueued on the stream before the call are visible to the peers' kernels enqueued after their call.  Every
 * rank must call it with the same, growing epoch sequence. */
/* Id exchange of the sharded path as peer-memory stores (replaces the count all-gather and the key all-to-all of
 * SURVEY 8(e) step (1); no reference counterpart).  dfm_shard_push_counts: this rank's per-owner unique-key counts (W)
 * become row `rank` of the W x W count matrix of every rank (peer_matrix[r] = rank r's matrix, int64).  After a
 * dfm_peer_barrier the matrix is complete everywhere; dfm_shard_push_keys then stores this rank's keys (send order,
 * grouped by owner) into each owner's receive buffer at the offset the matrix implies (receive order = grouped by
 * source rank); entries that would fall beyond `capacity` are dropped (the host sees the overflow in the matrix). */
DFM_API int dfm_shard_push_counts(const int64_t* counts, int world, int rank, int64_t* const* peer_matrix, void* stream);
DFM_API int dfm_shard_push_keys(const uint32_t* send_keys, const int64_t* matrix, int world, int rank, int64_t capacity,
                                uint32_t* const* peer_keys, void* stream);
DFM_API int dfm_peer_barrier(uint32_t* const* peer_flags, int world, int rank, uint32_t epoch, void* stream);
/* Per-field table source of a plan: row_stride / w1_stride (floats, 0 = dim / 1) let K1 read the field's rows
 * out of a strided buffer (the received reply rows); foreign = 1 marks a table whose gradient is produced
 * elsewhere: K1 emits no sort key for its ids and dfm_embed_bwd / dfm_rows_bwd ignore it.  On the sample-side
 * plan of the sharded path the SHARDED tables are foreign (the exchange kernels act on exactly those fields),
 * on the owner-side plan the REPLICATED (small) tables are.  dfm_plan_set_table_stride applies one stride
 * pair to every id table. */
DFM_API int dfm_plan_set_field_source(dfm_plan* plan, int field, int row_stride, int w1_stride, int foreign);
DFM_API int dfm_plan_set_table_stride(dfm_plan* plan, int row_stride, int w1_stride);
DFM_API size_t dfm_rows_bwd_workspace_bytes(const dfm_plan* plan, int64_t n_rows);
DFM_API int dfm_rows_bwd(const dfm_plan* plan, int64_t n_rows, const float* const* params,
                         const uint32_t* keys, const float* g_rows, const float* g_scalars, float l2,
                         const float* l2_gscale, int mode, float* const* grads, uint32_t* sorted_keys,
                         uint32_t* sorted_payload, float* row_grad2, float* row_grad1, int64_t* n_valid,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-sparse optimizer step on the id tables (reference trainer.py:232-237 +
 * trainer.py:67-78: clip_grad_norm_(1.0), torch.optim.Adam(lr 1e-3)).  Consumes dfm_embed_bwd's /
 * dfm_rows_bwd's row-sparse output in place (sorted keys + the summed gradient at the first position of
 * every segment): torch.optim.Adam's update restricted to the touched rows ("lazy" moments -- documented
 * deviation from the dense reference; oracle: adam_rows).  params / exp_avg / exp_avg_sq: 5 slots per field
 * like everywhere, only the two table slots are read.  clip_scale: device scalar multiplying every gradient
 * (min(1, max_norm / (norm + 1e-6)) of the global-norm clip), or NULL.
 *   ssq_acc (optional): device double holding sum ||w||^2 over the tables (dfm_sumsq_acc); the kernel adds
 *   sum(w_new^2 - w_old^2) of the touched elements (fp64, per-block partials added in block order) so the L2
 *   value needs no pass over the tables; ssq_workspace >= dfm_adam_rows_workspace_bytes() then.
 *   dfm_rows_sumsq: sum of squares of the touched rows' gradients (the table part of that norm; only the
 *   dim[f] defined floats of each row), deterministic fixed-order reduction into out[0].
 * ---------------------------------------------------------------------------------------- */
DFM_API int dfm_adam_rows(const dfm_plan* plan, int64_t n_sorted, const uint32_t* sorted_keys,
                          const float* row_grad2, const float* row_grad1, float* const* params,
                          float* const* exp_avg, float* const* exp_avg_sq, float lr, float beta1, float beta2,
                          float eps, int64_t step, const float* clip_scale, double* ssq_acc,
                          void* ssq_workspace, size_t ssq_workspace_bytes, void* stream);
DFM_API size_t dfm_adam_rows_workspace_bytes(void);
DFM_API size_t dfm_rows_sumsq_workspace_bytes(void);
DFM_API int dfm_rows_sumsq(const dfm_plan* plan, int64_t n_sorted, const uint32_t* sorted_keys,
                           const float* row_grad2, const float* row_grad1, float* out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * DNN tower (reference deepfm/models/layers/dnn.py:45-59: Linear -> BatchNorm1d ->
 * activation -> Dropout, and the head nn.Linear(., 1) of deepfm.py:36-42 / xdeepfm.py:42-48).
 *   dfm_gemm3: fp32-accurate GEMM on tcgen05 ("3xTF32": hi*hi + hi*lo + lo*hi of the tf32 split of both
 *       operands, fp32 accumulation in tensor memory; operands by TMA, no transposed copies):
 *         mode 0  D[M][N] = A[M][K] * B[N][K]^T (+ bias[N])     nn.Linear forward       (ATen addmm)
 *         mode 1  D[M][N] = A[M][K] * B[K][N]                   grad wrt the input       (ATen mm)
 *         mode 2  D[M][N] = A[K][M]^T * B[K][N]                 grad wrt the weight      (ATen mm), split-K with a
 *                 fixed-order reduction: workspace >= dfm_gemm3_workspace_bytes().  Internally the product runs
 *                 swapped (D^T = B^T A, tiles stored transposed) so that the lo part of the SMALL operand A is
 *                 precomputed like a weight's; same result, same interface.
 *       Every mode needs workspace >= dfm_gemm3_workspace_bytes() (lo part of the B-side operand, split-K partials).
 *       The contiguous extent of each operand must be a multiple of 4 floats and 16-byte aligned
 *       (DFM_ERR_UNSUPPORTED otherwise).
 *   dfm_bn_stats: training-mode BatchNorm1d statistics of y (M, C): mean, rstd = 1/sqrt(biased var + eps);
 *       running_mean / running_var (optional) get the momentum update with the unbiased variance.
 *   dfm_bn_act_fwd: out = dropout(act(gamma * (y - mean) * rstd + beta)); bn 0 = none, 1 = batch statistics,
 *       2 = fixed (running) statistics; act 0 relu, 1 leaky_relu(0.01), 2 gelu (erf), 3 tanh (dnn.py:20-25);
 *       the dropout mask is a counter-based function of (seed, element index), regenerated by the backward.
 *   dfm_bn_act_bwd: dy from da (z, xhat and the mask recomputed from y); dgamma, dbeta (BatchNorm affine grads),
 *       dbias (optional: column sums of dy, the Linear bias gradient).  workspace >= dfm_tower_workspace_bytes(M, C).
 *   dfm_head_fwd / dfm_head_bwd: out[m] = a[m, :] . w + b;  da (optional), dw (C), db (1).
 * All reductions are fixed-order with fp64 partials: deterministic.
 * ---------------------------------------------------------------------------------------- */
DFM_API size_t dfm_gemm3_workspace_bytes(int mode, int64_t M, int64_t N, int64_t K);
DFM_API int dfm_gemm3(int mode, const float* A, const float* B, float* D, const float* bias, int64_t M, int64_t N,
                      int64_t K, void* workspace, size_t workspace_bytes, void* stream);
DFM_API size_t dfm_tower_workspace_bytes(int64_t M, int C);
DFM_API int dfm_bn_stats(const float* y, int64_t M, int C, float eps, float* mean, float* rstd, float* running_mean,
                         float* running_var, float momentum, void* workspace, size_t workspace_bytes, void* stream);
DFM_API int dfm_bn_act_fwd(const float* y, int64_t M, int C, int bn, int act, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, float drop_p, uint64_t seed, float* out, void* stream);
DFM_API int dfm_bn_act_bwd(const float* da, const float* y, int64_t M, int C, int bn, int act, const float* mean,
                           const float* rstd, const float* gamma, const float* beta, float drop_p, uint64_t seed,
                           float* dy, float* dgamma, float* dbeta, float* dbias, void* workspace,
                           size_t workspace_bytes, void* stream);
/* The whole tower in two calls (same arithmetic: they sequence dfm_gemm3 / dfm_bn_stats / dfm_bn_act_fwd / _bwd block by
 * block on `stream`; one autograd node per tower instead of one per block -- the host side of a multi-rank step was the
 * bottleneck).  dims: n_layers + 1 widths (input, h_1 .. h_n).  params: per block W (h, din), bias, gamma, beta (NULL where
 * absent).  bn per block: 0 none, 1 batch statistics (running: running_mean / running_var updated with momentum[l], or NULL),
 * 2 fixed statistics (forward: running = (mean, rstd) pair; backward: fixed = the same pair).  store: the activations the
 * backward needs (dfm_tower_store_floats floats: pre-activations, block outputs, batch statistics).  grads: per block
 * dW, dbias, dgamma, dbeta (NULL where absent); dx optional. */
DFM_API size_t dfm_tower_store_floats(int n_layers, const int64_t* dims, int64_t M, int64_t* offsets);
DFM_API size_t dfm_tower_seq_workspace_bytes(int n_layers, const int64_t* dims, int64_t M, int backward);
DFM_API int dfm_tower_fwd(int n_layers, const int64_t* dims, int64_t M, const float* x, const float* const* params,
                          const int* bn, float* const* running, const float* eps, const float* momentum, int act,
                          float drop_p, const uint64_t* seeds, float* store, float* out, void* workspace,
                          size_t workspace_bytes, void* stream);
DFM_API int dfm_tower_bwd(int n_layers, const int64_t* dims, int64_t M, const float* x, const float* const* params,
                          const int* bn, const float* const* fixed, int act, float drop_p, const uint64_t* seeds,
                          const float* store, const float* g_out, float* dx, float* const* grads, void* workspace,
                          size_t workspace_bytes, void* stream);
DFM_API int dfm_head_fwd(const float* a, const float* w, const float* b, int64_t M, int C, float* out, void* stream);
DFM_API int dfm_head_bwd(const float* a, const float* w, const float* g, int64_t M, int C, float* da, float* dw, float* db,
                         void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPFM_B200_H */
