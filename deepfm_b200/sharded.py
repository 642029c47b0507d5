"""Row-sharded ``FeatureEmbedding`` for one 8xB200 box (the reference is single-device).

Every embedding table (second- and first-order) is sharded by row over the ``W`` ranks of a
``torch.distributed`` NCCL group: ``owner(f, id) = (id + f) mod W`` with ``f`` the schema index of the field,
``local_row = id div W`` (the rotation spreads the hot ids 1, 2, ... of the 26 tables over the ranks instead of
piling them on ranks 1, 2, ...; bit-exact contract: ``oracle.shard_route``).  Samples stay data-parallel:
rank ``r`` owns its ``b`` samples.  One step is

    forward   route ids by owner -> all-to-all(keys) -> owners gather rows (dfm_shard_gather)
              -> all-to-all(vectors, first-order weights) back -> fused gather/FM kernel K1 reads the
              received rows instead of a table (dfm_embed_fwd on a "received rows" plan)
    backward  dfm_shard_pack_grad builds the per-row gradient (FM backward fused) in send order
              -> all-to-all to the owners -> sorted segmented reduction on the owner (dfm_rows_bwd),
              so each row's gradient is produced on exactly one GPU: no allreduce of table gradients.

Dense-field Linears and everything else (DNN, CIN, attention) are replicated and data-parallel;
``allreduce_dense`` averages their gradients in one flat bucket.  Supported here: SPARSE, DENSE and
multi-hot SEQUENCE (sum / mean) fields with ``embedding_dim == fm_embed_dim`` (the Criteo shapes).
A bag is exchanged id by id (padding entries are not sent), pooled by K1 on the sample's GPU with the
bag's global non-pad count as the mean divisor, and every member's gradient row travels back to the
row's owner, so the owner-side reduction is unchanged.

Small tables (``vocabulary_size <= replicate_below``, default 4096 rows = 1 MB at D = 64) are REPLICATED, not
sharded (sharding them is pure overhead): every rank looks them up locally inside the same K1
launch, produces their dense gradient with the local sort/segment-reduce, and the reducer averages it with the
other data-parallel parameters.  On the Criteo shape that takes 14 of the 26 tables -- and the hottest ids of
the batch -- out of the exchange.

The phases are plain methods so that a single process can emulate ``W`` ranks (tests) and so that
the routing logic (pure torch ops) runs on CPU tensors under gloo.
"""

from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from .layers._status import IndexStatusMixin
from .layers.embedding import RowSparseGrads
from .schema import kind_of

VIRTUAL_VOCAB = 1 << 26     # capacity of the "received rows" pseudo table of the sample-side plan


def local_rows(vocab: int, world: int, rank: int = 0, field: int = 0) -> int:
    """Rows of a sharded table on EVERY rank: ceil(vocab / world).  The ids a rank owns in one table are one residue
    class, (id + field) mod world == rank, with local_row = id div world; sizing every shard to the largest class
    (at most one unused row) makes a row's local sort key the same number on every rank."""
    return max((vocab + world - 1) // world, 1)


def owned_rows(vocab: int, world: int, rank: int, field: int = 0) -> int:
    """Number of ids in [0, vocab) that ``rank`` owns in the table of schema field ``field``."""
    first = (rank - field) % world
    return max((vocab - first + world - 1) // world, 0)


@dataclass
class Route:
    """One batch routed for the unique-row exchange (see ShardedFeatureEmbedding)."""
    send_keys: torch.Tensor      # (>= n_unique,) int32: owner-local keys (vbase[f] + id div W) in send order
    counts: torch.Tensor         # (W,) int64: unique keys per destination
    pos: torch.Tensor            # (b * S_all,) int64, the plan's field-major blocks: 1 + index of the slot's key in the
    #                              send order, 0 = nothing sent (padding id of a bag, replicated table)
    skeys: Optional[torch.Tensor] = None   # kernel path: the sorted (owner << lbits | key, payload) stream and the
    spay: Optional[torch.Tensor] = None    # unique index of every sorted position -- the backward's segments
    uidx: Optional[torch.Tensor] = None
    n_sorted: int = 0


def field_positions(pos: torch.Tensor, b: int, lens: Sequence[int]) -> List[torch.Tensor]:
    """Per table field: its (b,) / (b, L) block of ``Route.pos`` (views, no copy) -- K1's id columns."""
    out, s0 = [], 0
    for L in lens:
        blk = pos[b * s0: b * (s0 + L)]
        out.append(blk.view(b, L) if L > 1 else blk)
        s0 += L
    return out


def route_unique(ids: torch.Tensor, vbase: torch.Tensor, world: int, lens: Optional[Sequence[int]] = None,
                 bag: Optional[Sequence[bool]] = None, rot: Optional[Sequence[int]] = None, lbits: int = 24) -> Route:
    """Torch restatement of the routing kernels (CPU and CUDA tensors; bit-exact contract: oracle.shard_route_unique).
    ids (b, S) int64 (the sharded id slots of one sample side by side, bags expanded), vbase (S,) int64 per slot,
    lens / bag / rot per table field.  key = owner << lbits | (vbase + id div W) with owner = (id + rot) mod W;
    padding entries (id 0) of bag fields are not sent; the send order is the sorted set of distinct keys."""
    b, S = ids.shape
    lens = [1] * S if lens is None else list(lens)
    bag = [False] * len(lens) if bag is None else list(bag)
    rot = [0] * len(lens) if rot is None else list(rot)
    slot_bag = torch.tensor([g for L, g in zip(lens, bag) for _ in range(L)], dtype=torch.bool, device=ids.device)
    slot_rot = torch.tensor([r for L, r in zip(lens, rot) for _ in range(L)], dtype=ids.dtype, device=ids.device)
    sent = ~(slot_bag[None, :] & (ids == 0))
    owner = (ids + slot_rot[None, :]) % world
    comp = (owner << lbits) | (vbase[None, :] + ids // world)
    flat, sent_f = comp.reshape(-1), sent.reshape(-1)
    uniq, inverse = torch.unique(flat[sent_f], sorted=True, return_inverse=True)
    counts = torch.bincount(uniq >> lbits, minlength=world)[:world]
    pos_flat = torch.zeros(b * S, dtype=torch.int64, device=ids.device)
    pos_flat[sent_f] = inverse + 1
    pos_bs = pos_flat.view(b, S)
    blocks, s0 = [], 0
    for L in lens:
        blocks.append(pos_bs[:, s0:s0 + L].reshape(-1))
        s0 += L
    return Route(send_keys=(uniq & ((1 << lbits) - 1)).to(torch.int32), counts=counts, pos=torch.cat(blocks))


class TorchDistComm:
    """all-to-all over the default (NCCL or gloo) process group; sizes travel through the host."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def exchange_counts(self, counts: torch.Tensor):
        """(send_counts, recv_counts) as host lists: ONE all-gather of the W x W count matrix, one sync."""
        rows = [torch.empty_like(counts) for _ in range(self.world)]
        self.dist.all_gather(rows, counts.contiguous(), group=self.group)
        host = torch.stack(rows).tolist()
        self.last_matrix = host
        return host[self.rank], [host[src][self.rank] for src in range(self.world)]

    def exchange_counts_async(self, counts: torch.Tensor):
        """Same exchange, but the W x W matrix is copied to pinned host memory without blocking; call
        ``finish_counts`` later (normally the copy finished long before)."""
        if counts.is_cuda:
            mat = torch.empty((self.world, counts.numel()), dtype=counts.dtype, device=counts.device)
            self.dist.all_gather_into_tensor(mat, counts.contiguous(), group=self.group)
            ring = self.__dict__.get("_pinned_ring")
            if ring is None:          # all four up front: a pinned allocation synchronises the device
                ring = self._pinned_ring = [torch.empty(mat.shape, dtype=mat.dtype, pin_memory=True) for _ in range(4)]
                self._pinned_i = 0
            host = ring[self._pinned_i % len(ring)]
            self._pinned_i += 1
        else:
            rows = [torch.empty_like(counts) for _ in range(self.world)]
            self.dist.all_gather(rows, counts.contiguous(), group=self.group)
            mat = torch.stack(rows)
            host = torch.empty(mat.shape, dtype=mat.dtype)
        host.copy_(mat, non_blocking=True)
        ev = torch.cuda.Event() if mat.is_cuda else None
        if ev is not None:
            ev.record()
        return host, ev

    def finish_counts(self, pending):
        host, ev = pending
        if ev is not None:
            ev.synchronize()
        mat = host.tolist()
        self.last_matrix = mat            # mat[s][r]: rows rank s sends to rank r
        return mat[self.rank], [mat[src][self.rank] for src in range(self.world)]

    def all_to_all(self, send: torch.Tensor, send_counts: Sequence[int], recv_counts: Sequence[int],
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
        self.dist.all_to_all_single(out, send.contiguous(), list(recv_counts), list(send_counts), group=self.group)
        return out


class PeerExchange:
    """The two vector exchanges of a step as peer-memory writes (compute fused with its collective).

    Every rank owns one symmetric-memory allocation (``deepfm_b200._peer``: CUDA VMM memory mapped into every peer
    over NVLink / NVSwitch) holding, for each direction (0: looked-up rows "got", 1: gradient rows "g_recv") and each
    step parity, a (cap, D) vector buffer followed by a (cap, 4) scalar buffer -- 256-byte rows that start on a
    256-byte boundary and 16-byte scalar records, so every peer store is a whole number of aligned sectors.  The
    owner-side gather kernel stores each reply row straight into the requesting GPU's ``got`` buffers, the sample-side
    segmented reduction stores each unique row's gradient straight into the owning GPU's ``g_recv`` buffers: no staging
    buffer, no NCCL all-to-all; one device-side barrier separates the writes from their consumer.  Buffers alternate
    between steps (a peer can be at most one exchange ahead)."""

    def __init__(self, comm: "TorchDistComm", dim: int, capacity_rows: int, device):
        from . import _peer
        self.comm, self.world, self.rank = comm, comm.world, comm.rank
        self.dim, self.cap = int(dim), (int(capacity_rows) + 63) // 64 * 64      # every block starts on a 256-byte boundary
        group = comm.group if comm.group is not None else comm.dist.group.WORLD
        n = 4 * self.cap * (self.dim + 4)
        W = self.world
        # tail of the allocation: [64 flag words: 4 barrier channels x 16 ranks][2 count matrices W x W int64][2 key
        # receive buffers of cap int32] -- the id exchange of the prefetch (dfm_shard_push_counts / _keys)
        self._off_flags = n
        self._off_matrix = n + 64
        self._mat_words = (2 * W * W + 3) // 4 * 4
        self._off_keys = self._off_matrix + 2 * self._mat_words
        total = self._off_keys + 2 * self.cap
        self.buf, self.hdl = _peer.symmetric_empty(total, device, group)
        self.buf.zero_()                                     # row 0 of both got buffers is the reserved zero row
        self.peer_base = [int(p) for p in self.hdl.buffer_ptrs]
        self._flag_ptrs = [_lib.ptr_array([b + 4 * (n + 16 * ch) for b in self.peer_base]) for ch in range(4)]
        self._epoch = [0, 0, 0, 0]
        self.step = 0
        self.pf_step = 0                                     # prefetches issued (parity of the id-exchange buffers)
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)                          # setup only: every rank's flags are zero before any epoch

    def _offsets(self, kind: int, parity: int):
        """Float offsets of the (vector, scalar) buffers of (kind, parity) inside a rank's allocation."""
        base = (2 * kind + parity) * self.cap * (self.dim + 4)
        return base, base + self.cap * self.dim

    def region(self, kind: int, parity: int, rank: Optional[int] = None):
        """Device addresses (vector buffer, scalar buffer) of (kind 0: got, 1: g_recv; parity) on ``rank``."""
        b = self.peer_base[self.rank if rank is None else rank]
        ov, os_ = self._offsets(kind, parity)
        return b + 4 * ov, b + 4 * os_

    def views(self, kind: int, parity: int, rows: int):
        ov, os_ = self._offsets(kind, parity)
        return (self.buf[ov: ov + rows * self.dim].view(rows, self.dim),
                self.buf[os_: os_ + rows * 4].view(rows, 4))

    def fits(self, matrix) -> bool:
        """Same answer on every rank (all of them hold the full W x W count matrix)."""
        W = self.world
        sent = max(sum(matrix[s]) for s in range(W))
        recv = max(sum(matrix[s][r] for s in range(W)) for r in range(W))
        return sent + 1 <= self.cap and recv <= self.cap

    def barrier(self, channel: int = 0) -> None:
        """All-ranks barrier on the current stream: one 32-thread kernel of our own over peer-mapped flag words
        (``dfm_peer_barrier``).  torch's ``handle.barrier`` costs ~1 ms of HOST time per call (measured with cProfile
        at W = 2: 1.9 of the 5.9 ms step), which made the sharded step host-bound.  Barriers issued on different
        streams use different channels (own flag words, own epoch): channels 0 / 1 the step's stream (forward,
        backward), channel 2 the prefetch stream."""
        self._epoch[channel] += 1
        _lib.check(_lib.lib().dfm_peer_barrier(self._flag_ptrs[channel], self.world, self.rank,
                                               self._epoch[channel] & 0xFFFFFFFF, _lib.stream_ptr()), "dfm_peer_barrier")

    def matrix_ptrs(self, parity: int):
        return _lib.ptr_array([b + 4 * (self._off_matrix + parity * self._mat_words) for b in self.peer_base])

    def matrix_view(self, parity: int) -> torch.Tensor:
        o = self._off_matrix + parity * self._mat_words
        return self.buf[o: o + 2 * self.world * self.world].view(torch.int64).view(self.world, self.world)

    def key_ptrs(self, parity: int):
        return _lib.ptr_array([b + 4 * (self._off_keys + parity * self.cap) for b in self.peer_base])

    def keys_view(self, parity: int, n: int) -> torch.Tensor:
        o = self._off_keys + parity * self.cap
        return self.buf[o: o + n].view(torch.int32)

    def push_ids(self, route: "Route"):
        """Prefetch-stream half of the id exchange: counts -> every rank's matrix, barrier, keys -> their owners' receive
        buffers, barrier; returns (pinned host matrix, event, parity).  No NCCL call, ~5 kernel launches."""
        lib = _lib.lib()
        p = self.pf_step & 1
        self.pf_step += 1
        st = _lib.stream_ptr()
        _lib.check(lib.dfm_shard_push_counts(_lib.ptr(route.counts), self.world, self.rank, self.matrix_ptrs(p), st),
                   "dfm_shard_push_counts")
        self.barrier(2)
        mat = self.matrix_view(p)
        _lib.check(lib.dfm_shard_push_keys(_lib.ptr(route.send_keys), mat.data_ptr(), self.world, self.rank, self.cap,
                                           self.key_ptrs(p), st), "dfm_shard_push_keys")
        self.barrier(2)
        ring = self.__dict__.get("_pinned_ring")
        if ring is None:
            ring = self._pinned_ring = [torch.empty((self.world, self.world), dtype=torch.int64, pin_memory=True) for _ in range(4)]
            self._pinned_i = 0
        host = ring[self._pinned_i % 4]
        self._pinned_i += 1
        host.copy_(mat, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev, p


class _ShardedEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "ShardedFeatureEmbedding", n_inputs: int, need_bwd: bool, *tensors):
        inputs, params = tensors[:n_inputs], tensors[n_inputs:]
        comm = mod.comm
        pref = mod._take_prefetch(inputs)
        pushed = None
        if pref is not None:      # routed ahead of time (prefetch): the counts are already on the host
            route, pending = pref
            if len(pending) == 3:                 # ids exchanged by peer stores: (host matrix, event, parity)
                pushed = pending[2]
                pending = pending[:2]
            send_counts, recv_counts = comm.finish_counts(pending)
        else:
            route = mod.route(inputs)
            send_counts, recv_counts = comm.exchange_counts(route.counts)     # the only host sync of the step
        n_send, n_recv = int(sum(send_counts)), int(sum(recv_counts))
        px = mod.peer_exchange(inputs[0].device)
        matrix = getattr(comm, "last_matrix", None)
        use_p2p = px is not None and matrix is not None and px.fits(matrix)
        if pushed is not None and use_p2p:
            recv_keys = px.keys_view(pushed, n_recv)          # landed before the prefetch event this stream waited on
        else:
            recv_keys = comm.all_to_all(route.send_keys[:n_send], send_counts, recv_counts)
        parity = 0
        if use_p2p:
            # owners store every reply row straight into the requester's got buffers (NVLink P2P), then one barrier
            parity = px.step & 1
            px.step += 1
            bkeys = mod.gather(recv_keys, p2p=(px, matrix, parity))
            px.barrier(0)
            got_vec, got_sc = px.views(0, parity, n_send + 1)
        else:
            vec, sc, bkeys = mod.gather(recv_keys)
            got_vec, got_sc = mod.reply_buffers(n_send, vec)                  # row 0: the reserved zero row
            comm.all_to_all(vec, recv_counts, send_counts, out=got_vec[1:])
            comm.all_to_all(sc, recv_counts, send_counts, out=got_sc[1:])
        bsorted = mod.sort_owner_keys(bkeys) if need_bwd else None            # side stream: the owner-side backward's sort
        first, field, flat, fm, fm_sum, aux, fin_inputs, keys = mod.finish(inputs, route.pos, got_vec, got_sc, need_bwd)
        mod._run_queued_prefetch()      # the NEXT batch's routing: behind this batch's exchange on the NCCL stream
        ctx.mod, ctx.n_inputs = mod, n_inputs
        ctx.counts = (send_counts, recv_counts)
        ctx.p2p = (px, matrix, parity) if use_p2p else None
        ctx.route = route if need_bwd else None
        ctx.bsorted = bsorted
        ctx.set_materialize_grads(False)
        ctx.l2 = None
        if need_bwd:
            # outputs go through save_for_backward (no reference cycle); the id / position tensors and the parameters
            # are inputs of this node and stay plain references (~170 tensors: packing / unpacking them costs ~0.2 ms)
            ctx.save_for_backward(field, flat, fm_sum, got_vec, got_sc, aux)
            ctx.fin_inputs, ctx.params = fin_inputs, params
            ctx.keys = keys
            mod._live_ctx = weakref.ref(ctx)
        anchor = torch.zeros((), device=first.device, dtype=torch.float32)     # see layers/l2.py
        return first, field, flat, fm, anchor

    @staticmethod
    def backward(ctx, g_first, g_field, g_flat, g_fm, _g_anchor=None):
        mod: ShardedFeatureEmbedding = ctx.mod
        mod._live_anchor = None
        cb = mod.__dict__.get("on_backward_start")
        if cb is not None:        # e.g. DenseGradReducer: every parameter downstream of the embedding has its gradient now
            cb()
        field, flat, fm_sum, got_vec, got_sc, aux = ctx.saved_tensors
        fin_inputs, params = ctx.fin_inputs, ctx.params
        send_counts, recv_counts = ctx.counts
        lam, gscale = ctx.l2 if ctx.l2 is not None else (0.0, None)
        ctx.l2 = None                     # consumed (see layers/l2.py)
        mod.raise_if_bad_index(block=False)
        cont = lambda g: None if g is None else g.contiguous()
        g_first, g_field, g_flat, g_fm = cont(g_first), cont(g_field), cont(g_flat), cont(g_fm)
        n_recv = int(sum(recv_counts))
        if ctx.p2p is not None:           # the unique rows' gradients are stored straight into the owners' buffers
            px, matrix, parity = ctx.p2p
            mod.reduce_grads(ctx.route, g_first, g_field, g_flat, g_fm, field, fm_sum, aux, p2p=ctx.p2p)
            dense_grads = mod.local_grads(fin_inputs, got_vec, got_sc, g_first, g_field, g_flat, g_fm, field, flat, fm_sum,
                                          params, lam, gscale, aux, ctx.keys)
            px.barrier(1)
            g_vec, g_sc = px.views(1, parity, n_recv)
        else:
            sv, ss = mod.reduce_grads(ctx.route, g_first, g_field, g_flat, g_fm, field, fm_sum, aux)
            dense_grads = mod.local_grads(fin_inputs, got_vec, got_sc, g_first, g_field, g_flat, g_fm, field, flat, fm_sum,
                                          params, lam, gscale, aux, ctx.keys)
            n_send = int(sum(send_counts))
            g_vec = mod.comm.all_to_all(sv[:n_send], send_counts, recv_counts)
            g_sc = mod.comm.all_to_all(ss[:n_send], send_counts, recv_counts)
        table_grads = mod.owner_backward(ctx.bsorted, g_vec, g_sc, params, lam, gscale)
        grads = [dense_grads.get(i, table_grads.get(i)) for i in range(len(params))]
        return (None, None, None) + (None,) * ctx.n_inputs + tuple(grads)


class ShardedFeatureEmbedding(IndexStatusMixin, nn.Module):
    """``grad_scale`` (default 1 / world): factor applied to every exchanged table-gradient row.  Each rank
    back-propagates the mean loss of ITS ``b`` samples; the data-parallel parameters' gradients are averaged over the
    ranks (``DenseGradReducer``), i.e. they are gradients of the GLOBAL-mean loss.  A sharded row's gradient is the SUM
    over the ranks of what their samples contribute, so it is scaled by 1 / world to be on that same scale (the global
    clip norm and the optimizer then see one consistent loss); the L2 term ``2*l2*w`` is added once by the owner,
    unscaled.  Set ``grad_scale = 1.0`` for sum semantics."""

    def __init__(self, schema, fm_embed_dim: int = 16, world: int = 1, rank: int = 0, comm=None,
                 replicate_below: int = 4096, grad_scale: Optional[float] = None) -> None:
        super().__init__()
        self._init_status()
        self.grad_scale = (1.0 / world) if grad_scale is None else float(grad_scale)
        self.schema, self.fm_embed_dim = schema, fm_embed_dim
        self.replicate_below = int(replicate_below)
        self.world, self.rank, self.comm = world, rank, comm
        self.field_names = list(schema.fields.keys())
        self.second_order_embeddings = nn.ModuleDict()
        self.first_order_embeddings = nn.ModuleDict()
        self.projections = nn.ModuleDict()            # always empty here (dims == fm_embed_dim)
        kinds, dims, vocabs, lvocabs, lens, combiners = [], [], [], [], [], []
        for name in self.field_names:
            fs = schema.fields[name]
            kind = kind_of(fs)
            comb = str(getattr(fs, "combiner", "mean") or "mean")
            if int(fs.embedding_dim) != fm_embed_dim or (kind == "sequence" and comb not in ("sum", "mean")):
                raise NotImplementedError(
                    f"field {name!r}: sharded tables support SPARSE / DENSE / sum- or mean-pooled SEQUENCE fields "
                    f"with embedding_dim == fm_embed_dim")
            d = int(fs.embedding_dim)
            if kind == "dense":
                self.second_order_embeddings[name] = nn.Linear(1, d)
                self.first_order_embeddings[name] = nn.Linear(1, 1)
                vocabs.append(0)
                lvocabs.append(0)
            else:
                replicated = int(fs.vocabulary_size) <= self.replicate_below
                rows = int(fs.vocabulary_size) if replicated else local_rows(int(fs.vocabulary_size), world)
                if kind == "sparse":
                    self.second_order_embeddings[name] = nn.Embedding(rows, d)
                    self.first_order_embeddings[name] = nn.Embedding(rows, 1)
                else:
                    self.second_order_embeddings[name] = nn.EmbeddingBag(rows, d, mode=comb)
                    self.first_order_embeddings[name] = nn.EmbeddingBag(rows, 1, mode=comb)
                vocabs.append(int(fs.vocabulary_size))
                lvocabs.append(rows)
            kinds.append(_lib.KIND[kind])
            dims.append(d)
            lens.append(int(fs.max_length) if kind == "sequence" else 1)
            combiners.append(_lib.COMBINER[comb] if kind == "sequence" else _lib.SUM)
        self._kinds, self._dims, self._vocabs, self._lvocabs = kinds, dims, vocabs, lvocabs
        self._max_lens, self._combiners = lens, combiners
        self.num_fields = len(kinds)
        self._repl_idx = [i for i, k in enumerate(kinds) if k != _lib.DENSE and vocabs[i] <= self.replicate_below]
        repl = set(self._repl_idx)
        self._table_idx = [i for i, k in enumerate(kinds) if k != _lib.DENSE and i not in repl]   # the SHARDED tables
        if not self._table_idx:
            raise NotImplementedError(
                f"no table has more than replicate_below={self.replicate_below} rows: nothing to shard -- use the plain "
                f"FeatureEmbedding on every rank and average all gradients (DenseGradReducer)")
        # (the exchange kernels need fm_embed_dim in {32, 64, 128} -- one warp load per row -- and say so when called)
        self._all_tables = [i for i, k in enumerate(kinds) if k != _lib.DENSE]
        self._all_lens = [lens[i] for i in self._all_tables]
        self._all_S = sum(self._all_lens)                                                          # id slots of the plan
        self._lens = [lens[i] for i in self._table_idx]
        self._bag = [kinds[i] == _lib.SEQUENCE for i in self._table_idx]
        self._S = sum(self._lens)
        self._A = sum(1 for i in range(len(kinds)) if kinds[i] == _lib.SEQUENCE and combiners[i] == _lib.MEAN)   # aux words / sample
        self._T = sum(dims)
        # owner-local key space: vbase[f] = rows of the sharded tables before field f (identical on every rank)
        vb, acc = [], 0
        for i in range(len(kinds)):
            vb.append(acc)
            if i in set(self._table_idx):
                acc += lvocabs[i]
        vb.append(acc)
        if acc >= 2 ** 31 or sum(vocabs) >= 2 ** 31:
            raise NotImplementedError("sharded tables: total rows must stay below 2^31")
        self._vbase, self._row_base = vb, vb                 # the owner-side plan's row_base is exactly this
        self._lbits = max(int(acc).bit_length(), 1)          # PAD of the local key space == acc must fit too
        self._kbits = self._lbits + int(world).bit_length()  # sort bits of owner << lbits | key (PAD = world << lbits)
        if self._kbits > 32:
            raise NotImplementedError("sharded tables: owner-major keys need more than 32 bits")
        self._pad_key = world << self._lbits
        self._max_tdim = fm_embed_dim
        self.grad_mode = "row_sparse"
        self.row_grads: Optional[RowSparseGrads] = None
        self.last_counts = None
        self._live_ctx = None
        self._live_anchor = None
        self._plans = None
        self._param_is_table: List[bool] = []
        self._slot_of_param: List[int] = []
        self._px = None                   # PeerExchange (None: not created yet, False: unavailable)
        self.p2p_capacity_rows = 0        # rows per exchange buffer; 0: sized by the first forward (b x slots + 16)
        self._side = None
        self._init_weights()

    def _init_weights(self) -> None:
        """Same family as the reference (embedding.py:66-74): xavier-uniform rows, zero padding row
        (global id 0 of field f lives on rank f mod W, local row 0), xavier Linears with zero bias."""
        index = {name: f for f, name in enumerate(self.field_names)}
        for name, m in list(self.second_order_embeddings.items()) + list(self.first_order_embeddings.items()):
            if isinstance(m, (nn.Embedding, nn.EmbeddingBag)):
                nn.init.xavier_uniform_(m.weight.data)
                if self.rank == index[name] % self.world or index[name] in self._repl_idx:
                    m.weight.data[0].zero_()
            else:
                nn.init.xavier_uniform_(m.weight.data)
                nn.init.zeros_(m.bias.data)

    @torch.no_grad()
    def load_from_full(self, full) -> None:
        """Take this rank's rows (id = (rank - f) mod W + W * local_row) and the replicated Linears from an
        unsharded FeatureEmbedding."""
        for f, name in enumerate(self.field_names):
            for mine, theirs in ((self.second_order_embeddings[name], full.second_order_embeddings[name]),
                                 (self.first_order_embeddings[name], full.first_order_embeddings[name])):
                if isinstance(mine, (nn.Embedding, nn.EmbeddingBag)) and f in self._repl_idx:
                    mine.weight.copy_(theirs.weight)
                elif isinstance(mine, (nn.Embedding, nn.EmbeddingBag)):
                    rows = theirs.weight[(self.rank - f) % self.world::self.world]
                    mine.weight[: rows.shape[0]].copy_(rows)
                    mine.weight[rows.shape[0]:].zero_()           # the (at most one) row no id maps to
                else:
                    mine.weight.copy_(theirs.weight)
                    mine.bias.copy_(theirs.bias)

    @torch.no_grad()
    def full_state_dict(self, group=None) -> Dict[str, torch.Tensor]:
        """Collective: the embedding's parameters in the REFERENCE layout (``FeatureEmbedding.state_dict`` keys and
        shapes: ``second_order_embeddings.<f>.weight (V, d)`` ...), every rank's rows interleaved back to
        ``id = (rank - f) mod W + W * local_row`` -- the inverse of ``load_from_full``.  Returned on every rank (CPU
        tensors), so rank 0 can ``torch.save`` a checkpoint the unsharded module / the reference loads
        (checkpoint interop for sharded tables)."""
        import torch.distributed as dist
        W = self.world
        out: Dict[str, torch.Tensor] = {}
        for f, name in enumerate(self.field_names):
            for prefix, mod in (("second_order_embeddings", self.second_order_embeddings[name]),
                                ("first_order_embeddings", self.first_order_embeddings[name])):
                if not isinstance(mod, (nn.Embedding, nn.EmbeddingBag)) or f in self._repl_idx:
                    for k, v in mod.state_dict().items():        # replicated: rank-local copy is the parameter
                        out[f"{prefix}.{name}.{k}"] = v.detach().cpu().clone()
                    continue
                V = self._vocabs[f]
                mine = mod.weight.detach()                        # (ceil(V / W), d) on every rank
                if W > 1:
                    parts = [torch.empty_like(mine) for _ in range(W)]
                    dist.all_gather(parts, mine.contiguous(), group=group)
                else:
                    parts = [mine]
                full = torch.empty((V, mine.shape[1]), dtype=mine.dtype)
                for r in range(W):
                    first = (r - f) % W                           # smallest id rank r owns in this table
                    n = len(range(first, V, W))
                    full[first::W] = parts[r][:n].cpu()
                out[f"{prefix}.{name}.weight"] = full
        return out

    # -- plans --------------------------------------------------------------------------------
    def _ensure_plans(self):
        if self._plans is None:
            lib = _lib.lib()
            n = self.num_fields

            def make(vocabs):
                p = lib.dfm_plan_create(n, _lib.i32_array(self._kinds), _lib.i32_array(self._dims),
                                        _lib.i64_array(vocabs), _lib.i32_array(self._max_lens),
                                        _lib.i32_array(self._combiners), int(self.fm_embed_dim))
                if not p:
                    raise ValueError(f"dfm_plan_create: {_lib.last_error()}")
                return C.c_void_p(p)
            self._virtual_cap = min(VIRTUAL_VOCAB, (2 ** 32 - 2) // max(len(self._table_idx), 1))
            shard = set(self._table_idx)
            virt = [self._virtual_cap if i in shard else v for i, v in enumerate(self._vocabs)]
            sample_plan = make(virt)
            for i in self._table_idx:     # sample side: a sharded table = the received reply rows ((n, D) vectors, (n, 4)
                # scalars with the first-order weight in column 0); its gradient is made by the owner
                _lib.check(lib.dfm_plan_set_field_source(sample_plan, i, self.fm_embed_dim, 4, 1), "dfm_plan_set_field_source")
            local_plan = make(self._lvocabs)
            for i in self._repl_idx:      # owner side: replicated tables are none of its business
                _lib.check(lib.dfm_plan_set_field_source(local_plan, i, 0, 0, 1), "dfm_plan_set_field_source")
            self._plans = (local_plan, sample_plan)
        return self._plans

    def __del__(self):
        plans = getattr(self, "_plans", None)
        if plans:
            for p in plans:
                try:
                    _lib.lib().dfm_plan_destroy(p)
                except Exception:
                    pass

    def _apply(self, fn, *args, **kwargs):
        self._ordered_cache = None        # .to() / .cuda() may replace the Parameter objects
        self._l2_split = None
        self.__dict__.pop("_ptrs_cache", None)
        self.__dict__.pop("_prepared_cache", None)
        self.__dict__.pop("_dense_bucket", None)
        return super()._apply(fn, *args, **kwargs)

    def _ordered_params(self) -> List[torch.Tensor]:
        cached = getattr(self, "_ordered_cache", None)
        if cached is not None:
            return cached
        out, is_table, slots = [], [], []
        for f, name in enumerate(self.field_names):
            second, first = self.second_order_embeddings[name], self.first_order_embeddings[name]
            dense = self._kinds[f] == _lib.DENSE
            shard = not dense and f not in self._repl_idx      # replicated tables are data-parallel parameters
            entries = [(0, second.weight, shard)] + ([(1, second.bias, False)] if dense else []) + \
                      [(2, first.weight, shard)] + ([(3, first.bias, False)] if dense else [])
            for k, p, tab in entries:
                out.append(p)
                is_table.append(tab)
                slots.append(5 * f + k)
        self._param_is_table, self._slot_of_param = is_table, slots
        self._ordered_cache = out
        return out

    def _ptrs(self, tensors, override: Optional[Dict[int, torch.Tensor]] = None) -> C.Array:
        canon = self.__dict__.get("_ordered_cache")
        if canon is not None and len(tensors) == len(canon) and tensors[0] is canon[0] and tensors[-1] is canon[-1]:
            key = (canon[0].data_ptr(), canon[-1].data_ptr())            # the parameters themselves: pointers are stable
            base = self.__dict__.get("_ptrs_cache")
            if base is None or base[0] != key:
                arr = (C.c_void_p * (5 * self.num_fields))()
                for slot, t in zip(self._slot_of_param, canon):
                    arr[slot] = t.data_ptr()
                base = self.__dict__["_ptrs_cache"] = (key, arr)
            arr = type(base[1]).from_buffer_copy(base[1])
            for slot, t in (override or {}).items():
                arr[slot] = t.data_ptr()
            return arr
        arr = (C.c_void_p * (5 * self.num_fields))()
        for slot, t in zip(self._slot_of_param, tensors):
            arr[slot] = None if t is None else t.data_ptr()
        for slot, t in (override or {}).items():
            arr[slot] = t.data_ptr()
        return arr

    def _side_stream(self, dev):
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    # -- phases -------------------------------------------------------------------------------
    def route(self, inputs: Sequence[torch.Tensor]) -> Route:
        """Sample side, ids only: owner-major keys -> sort -> unique keys in send order, per-owner counts, the slots'
        positions in that order, and the sorted stream the backward reduces over."""
        if not inputs[self._table_idx[0]].is_cuda:
            return self.route_torch(inputs)
        lib = _lib.lib()
        _, sample_plan = self._ensure_plans()
        dev = inputs[0].device
        b, W = inputs[0].shape[0], self.world
        n = b * self._S
        keys = torch.empty((max(n, 1),), device=dev, dtype=torch.int32)
        pay = torch.empty((max(n, 1),), device=dev, dtype=torch.int32)
        pos = torch.zeros((b * self._all_S,), device=dev, dtype=torch.int64)       # replicated tables' blocks stay 0
        counts = torch.empty((W,), device=dev, dtype=torch.int64)
        _lib.check(lib.dfm_shard_ukeys(sample_plan, W, _lib.i64_array(self._vbase), _lib.i64_array(self._vocabs), self._lbits,
                                       b, _lib.ptr_array(inputs), _lib.ptr(keys), _lib.ptr(pay), _lib.ptr(pos),
                                       _lib.ptr(self._status_word(dev)), _lib.stream_ptr()), "dfm_shard_ukeys")
        self._post_status()
        skeys, spay = torch.empty_like(keys), torch.empty_like(pay)
        ws = torch.empty((max(lib.dfm_sort_pairs_workspace_bytes(n, self._kbits), lib.dfm_shard_unique_workspace_bytes(n), 16),),
                         device=dev, dtype=torch.uint8)
        _lib.check(lib.dfm_sort_pairs(n, self._kbits, _lib.ptr(keys), _lib.ptr(pay), _lib.ptr(skeys), _lib.ptr(spay),
                                      ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "dfm_sort_pairs")
        ukeys, uidx = torch.empty_like(keys), torch.empty_like(keys)
        _lib.check(lib.dfm_shard_unique(sample_plan, W, self._lbits, b, n, _lib.ptr(skeys), _lib.ptr(spay), _lib.ptr(ukeys),
                                        _lib.ptr(uidx), _lib.ptr(pos), _lib.ptr(counts), ws.data_ptr(), ws.numel(),
                                        _lib.stream_ptr()), "dfm_shard_unique")
        return Route(send_keys=ukeys, counts=counts, pos=pos, skeys=skeys, spay=spay, uidx=uidx, n_sorted=n)

    def route_torch(self, inputs: Sequence[torch.Tensor]) -> Route:
        """The same routing in plain torch ops (CPU tensors under gloo; the restatement the GPU tests compare with)."""
        b = inputs[0].shape[0]
        ids = torch.cat([inputs[i].view(b, -1) for i in self._table_idx], dim=1)
        vb = torch.tensor([self._vbase[i] for i, L in zip(self._table_idx, self._lens) for _ in range(L)],
                          dtype=torch.int64, device=ids.device)
        r = route_unique(ids, vb, self.world, self._lens, self._bag, [i % self.world for i in self._table_idx], self._lbits)
        # positions in the plan's layout: one block per table field (replicated tables: zeros, nothing is sent)
        full = torch.zeros(b * self._all_S, dtype=torch.int64, device=ids.device)
        mine = dict(zip(self._table_idx, field_positions(r.pos, b, self._lens)))
        for i, blk in zip(self._all_tables, field_positions(full, b, self._all_lens)):
            if i in mine:
                blk.copy_(mine[i])
        r.pos = full
        return r

    def sharded_positions(self, pos: torch.Tensor, b: int) -> List[torch.Tensor]:
        """The position blocks of the SHARDED table fields (views of ``Route.pos``), in field order."""
        shard = set(self._table_idx)
        return [blk for i, blk in zip(self._all_tables, field_positions(pos, b, self._all_lens)) if i in shard]

    def reply_buffers(self, n: int, like: torch.Tensor):
        """((1 + n, D) vectors, (1 + n, 4) scalars) K1 reads as its table: row 0 is the reserved zero row (positions
        are 1-based), rows 1.. receive the replies in send order."""
        cap = getattr(self, "_virtual_cap", VIRTUAL_VOCAB)
        if n + 1 > cap:
            raise NotImplementedError(f"sharded tables: {n} exchanged rows per step exceed the plan capacity {cap}")
        vec = like.new_empty((n + 1, self.fm_embed_dim))
        sc = like.new_empty((n + 1, 4))
        vec[0].zero_()
        sc[0].zero_()
        return vec, sc

    def peer_exchange(self, device):
        """The symmetric-memory exchange buffers (created on first use; None when peer memory is unavailable or
        disabled with DFM_SHARD_P2P=0 -- the NCCL all-to-all path is used then)."""
        import os
        if self._px is False or self.comm is None or self.world == 1 or os.environ.get("DFM_SHARD_P2P", "1") == "0":
            return None
        if self._px is None:
            try:
                cap = int(getattr(self, "p2p_capacity_rows", 0)) or None
                if cap is None:
                    return None           # sized by the first forward (needs the batch size)
                self._px = PeerExchange(self.comm, self.fm_embed_dim, cap, device)
            except Exception as e:        # no P2P / symmetric memory on this box: keep the NCCL path
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({e}); using NCCL all-to-all")
                self._px = False
                return None
        return self._px

    def gather(self, recv_keys: torch.Tensor, p2p=None):
        """Owner side: the received unique keys -> rows.  Without ``p2p``: returns staging (vectors (M, D), scalars
        (M, 4), backward keys (M,)); with ``p2p = (px, matrix, parity)`` every row is stored straight into the requesting
        GPU's got buffers and only the backward keys are returned."""
        lib = _lib.lib()
        local_plan, _ = self._ensure_plans()
        params = self._ordered_params()
        M, W, me = recv_keys.numel(), self.world, self.rank
        dev = params[0].device
        bkeys = torch.empty((max(M, 1),), device=dev, dtype=torch.int32)
        if p2p is None:
            vec = torch.empty((M, self.fm_embed_dim), device=dev, dtype=torch.float32)
            sc = torch.empty((M, 4), device=dev, dtype=torch.float32)
            starts, vecs, scs = [0, M], [vec.data_ptr()], [sc.data_ptr()]
        else:
            px, matrix, parity = p2p
            starts, vecs, scs, acc = [0], [], [], 0
            for s_ in range(W):                                   # received keys are grouped by source rank
                acc += matrix[s_][me]
                starts.append(acc)
                send_off = sum(matrix[s_][:me])                   # where rank me's segment starts in s_'s send order
                v, c = px.region(0, parity, s_)
                vecs.append(v + (1 + send_off) * self.fm_embed_dim * 4)
                scs.append(c + (1 + send_off) * 16)
        if M > 0:
            _lib.check(lib.dfm_shard_gather2(local_plan, W, me, M, _lib.ptr(recv_keys), self._ptrs(params), len(vecs),
                                             _lib.i64_array(starts), _lib.ptr_array(vecs), _lib.ptr_array(scs),
                                             _lib.ptr(bkeys), _lib.stream_ptr()), "dfm_shard_gather2")
        bkeys = bkeys[:M]
        return bkeys if p2p is not None else (vec, sc, bkeys)

    def sort_owner_keys(self, bkeys: torch.Tensor):
        """Owner side: sort the backward keys NOW on a side stream (they depend on the ids only); the owner-side
        backward then starts at the segmented reduction.  Returns (sorted_keys, sorted_payload, event, M)."""
        lib = _lib.lib()
        local_plan, _ = self._ensure_plans()
        M = bkeys.numel()
        dev = bkeys.device
        if M == 0:
            return None
        cur = torch.cuda.current_stream(dev)
        side = self._side_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            skeys = torch.empty((M,), device=dev, dtype=torch.int32)
            spay = torch.empty((M,), device=dev, dtype=torch.int32)
            pay = torch.arange(M, device=dev, dtype=torch.int32)       # payload = row of the received buffers
            ws = torch.empty((max(lib.dfm_sort_pairs_workspace_bytes(M, self._lbits), 16),), device=dev, dtype=torch.uint8)
            _lib.check(lib.dfm_sort_pairs(M, self._lbits, bkeys.data_ptr(), pay.data_ptr(), skeys.data_ptr(), spay.data_ptr(),
                                          ws.data_ptr(), ws.numel(), side.cuda_stream), "dfm_sort_pairs")
            ev = torch.cuda.Event()
            ev.record(side)
        bkeys.record_stream(side)
        return skeys, spay, ev, M

    def _virtual_ptrs(self, params, got_vec, got_sc):
        """Sample-side plan: every sharded id table is the received row buffers."""
        arr = self._ptrs(params)
        for i in self._table_idx:
            arr[5 * i + 0] = got_vec.data_ptr()
            arr[5 * i + 2] = got_sc.data_ptr()
        return arr

    def finish(self, inputs, pos, got_vec, got_sc, need_bwd: bool):
        lib = _lib.lib()
        _, sample_plan = self._ensure_plans()
        params = self._ordered_params()
        dev = got_vec.device
        b = inputs[0].shape[0]
        F, D, T = self.num_fields, self.fm_embed_dim, self._T
        blocks = dict(zip(self._table_idx, self.sharded_positions(pos, b)))
        fin_inputs = [blocks.get(i, inputs[i]) for i in range(F)]     # sharded: reply positions; replicated / DENSE: the input
        flat = torch.empty((b, T), device=dev, dtype=torch.float32)
        field = flat.view(b, F, D)
        first = torch.empty((b, 1), device=dev, dtype=torch.float32)
        fm = torch.empty((b, 1), device=dev, dtype=torch.float32)
        fm_sum = torch.empty((b, D), device=dev, dtype=torch.float32) if need_bwd else None
        aux = torch.empty((b, max(self._A, 1)), device=dev, dtype=torch.int32)
        # sort keys of the replicated tables' ids (sharded slots get the PAD key): their backward is local
        keys = torch.empty((b * self._all_S,), device=dev, dtype=torch.int32) if (need_bwd and self._repl_idx) else None
        _lib.check(lib.dfm_embed_fwd(sample_plan, b, _lib.ptr_array(fin_inputs), self._virtual_ptrs(params, got_vec, got_sc),
                                     first.data_ptr(), field.data_ptr(), flat.data_ptr(), fm.data_ptr(),
                                     _lib.ptr(fm_sum), _lib.ptr(keys), aux.data_ptr(), None, _lib.stream_ptr()), "dfm_embed_fwd")
        return first, field, flat, fm, fm_sum, aux, fin_inputs, keys

    def reduce_grads(self, route: Route, g_first, g_field, g_flat, g_fm, field, fm_sum, aux, p2p=None):
        """Sample side backward of the sharded tables: the gradient rows of all slots that share a key are summed (in
        the sorted order of ``route``) and one row per unique key goes to its owner -- straight into the owner's
        g_recv buffers with ``p2p = (px, matrix, parity)``, else into staging (vectors, scalars) in send order."""
        lib = _lib.lib()
        _, sample_plan = self._ensure_plans()
        dev = field.device
        b = field.shape[0]
        W, me = self.world, self.rank
        n = route.n_sorted
        if p2p is None:
            sv = torch.empty((max(n, 1), self.fm_embed_dim), device=dev, dtype=torch.float32)
            ss = torch.empty((max(n, 1), 4), device=dev, dtype=torch.float32)
            starts, vecs, scs = [0, n], [sv.data_ptr()], [ss.data_ptr()]
        else:
            px, matrix, parity = p2p
            sv = ss = None
            starts, vecs, scs, acc = [0], [], [], 0
            for r_ in range(W):                               # my send order is grouped by owner rank
                acc += matrix[me][r_]
                starts.append(acc)
                recv_off = sum(matrix[s_][r_] for s_ in range(me))   # where my segment starts in r_'s receive order
                v, c = px.region(1, parity, r_)
                vecs.append(v + recv_off * self.fm_embed_dim * 4)
                scs.append(c + recv_off * 16)
        if n > 0 and b > 0:
            ws = torch.empty((max(lib.dfm_rows_bwd_workspace_bytes(sample_plan, n), 16),), device=dev, dtype=torch.uint8)
            _lib.check(lib.dfm_shard_bwd_peer(sample_plan, b, _lib.ptr(g_first), _lib.ptr(g_field), _lib.ptr(g_flat),
                                              _lib.ptr(g_fm), field.data_ptr(), _lib.ptr(fm_sum), _lib.ptr(aux),
                                              _lib.ptr(route.skeys), _lib.ptr(route.spay), _lib.ptr(route.uidx), n,
                                              self._pad_key, len(vecs), _lib.i64_array(starts), _lib.ptr_array(vecs),
                                              _lib.ptr_array(scs), float(self.grad_scale), ws.data_ptr(), ws.numel(),
                                              _lib.stream_ptr()), "dfm_shard_bwd_peer")
        return sv, ss

    def local_grads(self, fin_inputs, got_vec, got_sc, g_first, g_field, g_flat, g_fm, field, flat, fm_sum,
                    params, lam, gscale, aux=None, keys=None):
        """Data-parallel parameters of the embedding: DENSE-field Linears and the replicated (small) tables.
        K2 on the sample-side plan; the sharded tables are foreign there, so only the replicated ones are
        sorted / segment-reduced (dense (V, d) gradients, 2*l2*w on every row like the reference)."""
        lib = _lib.lib()
        _, sample_plan = self._ensure_plans()
        self._ordered_params()
        dev = flat.device
        b = flat.shape[0]
        dense_grads: Dict[int, torch.Tensor] = {}
        # all data-parallel gradients of the embedding live in ONE flat buffer (one allocation instead of ~65, and the
        # reducer all-reduces the buffer itself: no torch.cat, no copy back)
        layout = self.__dict__.get("_dense_layout")
        if layout is None or layout[0] != len(params):
            items, off = [], 0
            for i, (p, tab) in enumerate(zip(params, self._param_is_table)):
                if not tab:
                    items.append((i, off, p.numel(), tuple(p.shape)))
                    off += (p.numel() + 3) // 4 * 4
            layout = self.__dict__["_dense_layout"] = (len(params), items, off)
        _, items, total = layout
        # The bucket is PERSISTENT (like DDP's gradient_as_bucket_view): allocated once, its ~65 per-parameter views built
        # once; every step K2 overwrites it in place.  Safe in stream order: the previous step's allreduce on it was waited
        # for in DenseGradReducer.finish() before this step's backward was enqueued.
        # (Gradient accumulation -- a second backward while the parameters still hold the previous gradients, which are
        # views of the bucket -- must not overwrite them: that pass gets a fresh buffer.)
        bucket = self.__dict__.get("_dense_bucket")
        accumulating = any(getattr(params[i], "grad", None) is not None for i, _, _, _ in items)
        if accumulating or bucket is None or bucket[0].numel() != max(total, 1) or bucket[0].device != dev:
            flat_g = torch.zeros((max(total, 1),), device=dev, dtype=torch.float32)   # zeros: alignment gaps stay finite
            views = [(i, flat_g[off:off + n].view(shape)) for i, off, n, shape in items]
            bucket = (flat_g, views)
            if not accumulating:
                self.__dict__["_dense_bucket"] = bucket
        flat_g, views = bucket
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        for i, g in views:
            grads[i] = g
            dense_grads[i] = g
        self.dense_grad_flat = flat_g if items else None
        if dense_grads:
            keys = keys if self._repl_idx else None
            mode = _lib.GRAD_DENSE if keys is not None else _lib.GRAD_SKIP_TABLES
            ws = torch.empty((max(lib.dfm_embed_bwd_workspace_bytes(sample_plan, b), 16),), device=dev, dtype=torch.uint8)
            skeys = spay = None
            if keys is not None:
                skeys = torch.empty_like(keys)
                spay = torch.empty_like(keys)
            _lib.check(lib.dfm_embed_bwd(
                sample_plan, b, _lib.ptr_array(fin_inputs), self._virtual_ptrs(params, got_vec, got_sc),
                _lib.ptr(g_first), _lib.ptr(g_field), _lib.ptr(g_flat), _lib.ptr(g_fm), field.data_ptr(), flat.data_ptr(),
                _lib.ptr(fm_sum), _lib.ptr(keys), _lib.ptr(aux), float(lam), _lib.ptr(gscale), mode, self._ptrs(grads),
                _lib.ptr(skeys), _lib.ptr(spay), None, None, None, ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "dfm_embed_bwd")
        return dense_grads

    def owner_backward(self, bsorted, g_vec, g_sc, params, lam, gscale):
        """Owner side: K2 on the received unique rows (keys sorted ahead by ``sort_owner_keys``); every row's gradient
        is produced on exactly one GPU."""
        lib = _lib.lib()
        local_plan, _ = self._ensure_plans()
        self._ordered_params()
        dev = g_vec.device
        rowsparse = self.grad_mode == "row_sparse"
        grads, table_grads = [], {}
        for i, (p, tab) in enumerate(zip(params, self._param_is_table)):
            g = torch.empty_like(p) if (tab and not rowsparse) else None
            grads.append(g)
            if g is not None:
                table_grads[i] = g
        M = 0 if bsorted is None else bsorted[3]
        counts = torch.zeros((2,), device=dev, dtype=torch.int64)
        if M == 0:
            for g in table_grads.values():
                g.zero_()
            self.row_grads = None
            self.last_counts = counts
            return table_grads
        skeys, spay, ev, _ = bsorted
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        skeys.record_stream(cur)
        spay.record_stream(cur)
        ws = torch.empty((max(lib.dfm_rows_bwd_workspace_bytes(local_plan, M), 16),), device=dev, dtype=torch.uint8)
        rg2 = rg1 = None
        if rowsparse:
            rg2 = torch.empty((M, self.fm_embed_dim), device=dev, dtype=torch.float32)
            rg1 = torch.empty((M,), device=dev, dtype=torch.float32)
        mode = (_lib.GRAD_ROWSPARSE if rowsparse else _lib.GRAD_DENSE) | _lib.GRAD_PRESORTED
        _lib.check(lib.dfm_rows_bwd(local_plan, M, self._ptrs(params), None, _lib.ptr(g_vec), _lib.ptr(g_sc),
                                    float(lam), _lib.ptr(gscale), mode, self._ptrs(grads), skeys.data_ptr(), spay.data_ptr(),
                                    _lib.ptr(rg2), _lib.ptr(rg1), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _lib.stream_ptr()), "dfm_rows_bwd")
        self.row_grads = RowSparseGrads(skeys, spay, rg2, rg1, counts, self._row_base, self._dims, self.field_names) \
            if rowsparse else None
        self.last_counts = counts
        return table_grads

    # -- input pipeline hook ------------------------------------------------------------------
    def prefetch(self, batch, ready_event=None) -> None:
        """Route a FUTURE batch now (routing depends on the ids only, not on the weights).  The routing kernels, the
        count exchange and the copy of the W x W counts to pinned host memory run on the module's side stream, so
        calling this at the START of a step (for the next batch) hides all of it -- including the host's wait for the
        counts -- under the current step; the forward of that batch then starts without a host sync.
        Every rank must call this at the same point of its step (it contains a collective)."""
        inputs = self._prepare(batch)
        dev = inputs[0].device
        cur = torch.cuda.current_stream(dev)
        side = self._side_stream(dev)
        side.wait_stream(cur)                                  # the batch is ready on the caller's stream ...
        if ready_event is not None:
            side.wait_event(ready_event)                       # ... or when this event (e.g. its host -> device copy) fires
        px = self._px if isinstance(self._px, PeerExchange) else None
        with torch.cuda.stream(side):
            route = self.route(inputs)
            if px is not None and route.skeys is not None and os.environ.get("DFM_SHARD_P2P_IDS", "1") != "0":
                pending = px.push_ids(route)             # counts + keys as peer stores: no NCCL call in the prefetch
            else:
                pending = self.comm.exchange_counts_async(route.counts)
            ev = torch.cuda.Event()
            ev.record(side)
        for t in inputs:
            t.record_stream(side)
        # keyed by the batch's tensors: the prefetch of batch i + 1 is issued at the START of step i, i.e. before the
        # forward of batch i has taken its own entry (a single slot would be overwritten every step)
        store = self.__dict__.setdefault("_prefetched", {})
        while len(store) >= 2:
            store.pop(next(iter(store)))
        store[tuple(t.data_ptr() for t in inputs)] = (None, inputs, route, pending, ev)

    def queue_prefetch(self, batch, ready_event=None) -> None:
        """Ask for ``prefetch(batch)`` to run inside the NEXT forward, right after that forward has enqueued its own
        exchange.  All of torch's NCCL collectives share one stream in issue order: a prefetch issued at the very start
        of a step puts the next batch's routing kernels + count all-gather IN FRONT of the current batch's key exchange
        and delays its gather / K1 by ~0.5 ms (measured at W = 2); issued here it hides under the DNN forward, and it
        is still complete long before the next step's forward asks for the counts (no host wait)."""
        self.__dict__["_queued_prefetch"] = (batch, ready_event)

    def _run_queued_prefetch(self) -> None:
        q = self.__dict__.pop("_queued_prefetch", None)
        if q is not None:
            self.prefetch(q[0], ready_event=q[1])

    def _take_prefetch(self, inputs):
        store = self.__dict__.get("_prefetched")
        pref = store.pop(tuple(t.data_ptr() for t in inputs), None) if store else None
        if pref is None:
            return None
        cur = torch.cuda.current_stream(inputs[0].device)
        cur.wait_event(pref[4])                                 # the routing tensors were produced on the side stream
        route = pref[2]
        for t in (route.send_keys, route.counts, route.pos, route.skeys, route.spay, route.uidx):
            if t is not None:
                t.record_stream(cur)
        return route, pref[3]

    # -- module API ---------------------------------------------------------------------------
    def _prepare(self, batch):
        # the same batch is prepared twice (prefetch, then forward): remember the last few by the identity of their tensors
        key = tuple(map(id, batch.values())) if isinstance(batch, dict) else None
        cache = self.__dict__.setdefault("_prepared_cache", {})
        hit = cache.get(key) if key is not None else None
        if hit is not None and all(a is b for a, b in zip(hit[0], batch.values())):
            return hit[1]
        out = self._prepare_uncached(batch)
        if key is not None:
            while len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[key] = (tuple(batch.values()), out)      # holds the tensors: an id cannot be recycled while cached
        return out

    def _prepare_uncached(self, batch):
        out = []
        for i, name in enumerate(self.field_names):
            x = _lib.require_cuda(batch[name], f"batch[{name!r}]")
            x = x.float() if self._kinds[i] == _lib.DENSE else x.long()
            if self._kinds[i] == _lib.SEQUENCE:
                if x.dim() != 2 or x.shape[1] != self._max_lens[i]:
                    raise ValueError(f"batch[{name!r}]: expected (B, {self._max_lens[i]}) ids, got {tuple(x.shape)}")
                out.append(x.contiguous())
            else:
                out.append(x.reshape(x.shape[0]).contiguous())
        return out

    def forward_fused(self, batch):
        if self.comm is None:
            raise RuntimeError("ShardedFeatureEmbedding needs a communicator (e.g. TorchDistComm())")
        self._ensure_plans()
        params = self._ordered_params()
        # the cached pointer array is validated against every parameter's storage once per forward (a `p.data = ...`
        # assignment between steps is the one way a pointer can change without going through _apply)
        sig = tuple(p.data_ptr() for p in params)
        if sig != self.__dict__.get("_ptr_sig"):
            self.__dict__["_ptr_sig"] = sig
            self.__dict__.pop("_ptrs_cache", None)
        inputs = self._prepare(batch)
        if not self.p2p_capacity_rows:    # every rank derives the same capacity (same batch size and schema)
            self.p2p_capacity_rows = inputs[0].shape[0] * max(self._S, 1) + 16
        need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if self._status_pending and self.check_indices:
            self.raise_if_bad_index(block=True, keep=2)
        first, field, flat, fm, anchor = _ShardedEmbedFn.apply(self, len(inputs), need_bwd, *inputs, *params)
        self._live_anchor = anchor if (need_bwd and anchor.requires_grad) else None
        field._dfm_fm = (fm, field._version)
        return first, field, flat, fm

    def forward(self, batch):
        first, field, flat, _ = self.forward_fused(batch)
        return first, field, flat

    def table_parameters(self):
        return [p for p, t in zip(self._ordered_params(), self._param_is_table) if t]


class DenseGradReducer:
    """Data-parallel averaging of the replicated parameters' gradients, overlapped with the table backward.

    ``early`` parameters (DNN / CIN / attention / head Linears) have their gradients before the embedding backward
    starts: the autograd engine runs AccumulateGrad nodes ahead of every other ready node (they carry the maximal
    sequence number), and every early parameter sits downstream of the embedding, so when the embedding's backward node
    starts all of them are accumulated.  With ``embedding=`` the sharded module calls ``launch`` at that moment (one
    callback per step); without it, per-parameter post-accumulate-grad hooks count the arrivals.  Either way ONE flat
    NCCL allreduce is launched asynchronously and runs on NCCL's stream underneath the gradient exchange and the
    owner-side reduction.  ``late`` parameters (the DENSE-field Linears and replicated tables inside the embedding)
    follow in a second allreduce in ``finish()`` -- directly on the embedding's flat gradient buffer when it has one."""

    def __init__(self, early, late, world: int, group=None, embedding=None):
        import torch.distributed as dist
        self.dist, self.group, self.world = dist, group, world
        self.early = [p for p in early if p.requires_grad]
        self.late = [p for p in late if p.requires_grad]
        self.embedding = embedding
        self._pending, self._work, self._flat = 0, None, None
        self._shapes = [(p.numel(), tuple(p.shape)) for p in self.early]
        if world > 1:
            if embedding is not None and hasattr(embedding, "forward_fused"):
                embedding.on_backward_start = self.launch
            else:
                for p in self.early:
                    p.register_post_accumulate_grad_hook(self._hook)

    def launch(self) -> None:
        if self._work is not None or any(p.grad is None for p in self.early):
            return                        # already launched / a parameter has no gradient yet: finish() handles it
        self._flat = torch.cat([p.grad.reshape(-1) for p in self.early])
        self._work = self.dist.all_reduce(self._flat, group=self.group, async_op=True)

    def _hook(self, _param) -> None:
        self._pending += 1
        if self._pending == len(self.early):
            self.launch()

    @staticmethod
    def _scatter(flat, params, world):
        """The averaged bucket becomes the parameters' gradients: views, no per-parameter copies."""
        flat.div_(world)
        off = 0
        for p in params:
            n = p.grad.numel()
            p.grad = flat[off:off + n].view_as(p.grad)
            off += n

    def finish(self) -> None:
        """Call after ``backward()``: waits for the early bucket, reduces the late one, writes both back."""
        if self.world == 1:
            return
        if self._work is None:            # a parameter got no gradient this step: reduce what there is, now
            got = [p for p in self.early if p.grad is not None]
            if got:
                flat = torch.cat([p.grad.reshape(-1) for p in got])
                self.dist.all_reduce(flat, group=self.group)
                self._scatter(flat, got, self.world)
        else:
            self._work.wait()
            self._scatter(self._flat, self.early, self.world)
        late = [p for p in self.late if p.grad is not None]
        buf = getattr(self.embedding, "dense_grad_flat", None) if self.embedding is not None else None
        if buf is not None and late and sum(p.grad.numel() for p in late) <= buf.numel() \
                and all(p.grad.untyped_storage().data_ptr() == buf.untyped_storage().data_ptr() for p in (late[0], late[-1])):
            self.dist.all_reduce(buf, group=self.group)      # the gradients ARE views of this buffer
            buf.div_(self.world)
        elif late:
            flat = torch.cat([p.grad.reshape(-1) for p in late])
            self.dist.all_reduce(flat, group=self.group)
            self._scatter(flat, late, self.world)
        self._pending, self._work, self._flat = 0, None, None


def allreduce_dense(params, world: int, group=None) -> None:
    """Average the gradients of the data-parallel (replicated) parameters in one flat NCCL allreduce."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or world == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
