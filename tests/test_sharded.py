"""Row-sharded embedding path.

CPU (gloo, world_size 2): the routing rule and the all-to-all plumbing, bit-exact vs oracle.shard_route.
GPU: W ranks emulated in one process (the exchange is done by slicing) -- forward views and every
gradient of the sharded module must equal the unsharded FeatureEmbedding on the concatenated batch."""

import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from deepfm_b200.schema import DatasetSchema, FeatureType, FieldSchema
from deepfm_b200.sharded import Route, field_positions, local_rows, owned_rows, route_unique
from oracle import deepfm_oracle as O
from tests.helpers import assert_close_rel


def test_route_unique_matches_oracle():
    rng = np.random.default_rng(0)
    b, S, W, lbits = 57, 5, 4, 12
    vocab = [11, 300, 7, 1000, 50]
    ids = np.stack([((rng.zipf(1.3, b) - 1) % v) for v in vocab], axis=1).astype(np.int64)      # repeats: the point of dedup
    lv = [local_rows(v, W) for v in vocab]
    vbase = np.concatenate([[0], np.cumsum(lv)[:-1]]).astype(np.int64)
    rot = np.arange(S) % W
    r = route_unique(torch.from_numpy(ids), torch.from_numpy(vbase), W, rot=list(rot), lbits=lbits)
    send, counts, pos, comp = O.shard_route_unique(ids.reshape(-1), np.tile(vbase, b), W, lbits, rot=np.tile(rot, b))
    assert r.counts.tolist() == counts.tolist() and r.send_keys.tolist() == send.tolist()
    assert np.array_equal(r.pos.numpy(), O.shard_positions_from(pos, b, [1] * S))
    assert len(send) < b * S                                   # duplicates travel once
    # every slot finds ITS row: the key at its position, owned by its owner
    offs = np.concatenate([[0], np.cumsum(counts)])
    owner_of_pos = np.searchsorted(offs, pos - 1, side="right") - 1
    assert np.array_equal(owner_of_pos, (ids.reshape(-1) + np.tile(rot, b)) % W)
    assert np.array_equal(send[pos - 1], np.tile(vbase, b) + ids.reshape(-1) // W)
    assert local_rows(10, 4) == 3 and local_rows(2, 4) == 1 and local_rows(1000, 8) == 125
    # rotated rule: field 1 puts id 0 on rank 1, so rank 0 owns ids 3 and 7 of a 10-row table
    assert owned_rows(10, 4, 0, field=1) == 2 and owned_rows(10, 4, 1, field=1) == 3
    assert sum(owned_rows(1000, 8, r, field=5) for r in range(8)) == 1000


def test_route_unique_multihot_skips_bag_padding():
    """Bags are exchanged id by id; their padding entries (id 0, anywhere in the bag) are not sent, while the
    padding id of a SPARSE field is (its row 0 is returned as stored, embedding.py:35-40)."""
    rng = np.random.default_rng(1)
    b, W, lbits = 41, 3, 10
    lens, bag, vocab = [1, 4, 1, 6], [False, True, False, True], [30, 17, 5, 200]
    cols = []
    for L, g, v in zip(lens, bag, vocab):
        x = rng.integers(0, v, (b, L))
        if g:
            x = x * (rng.random((b, L)) < 0.6)             # pads in the middle of the bag too
            x[3] = 0                                       # an all-pad bag
        cols.append(x)
    ids = np.concatenate(cols, axis=1).astype(np.int64)
    lv = [local_rows(v, W) for v in vocab]
    fbase = np.concatenate([[0], np.cumsum(lv)[:-1]])
    vbase = np.repeat(fbase, lens).astype(np.int64)
    slot_bag = np.repeat(bag, lens)
    sent = ~(slot_bag[None, :] & (ids == 0))
    rot = [2, 0, 1, 2]                                     # owner = (id + rot[field]) mod W
    slot_rot = np.repeat(rot, lens)
    r = route_unique(torch.from_numpy(ids), torch.from_numpy(vbase), W, lens, bag, rot, lbits)
    send, counts, pos, comp = O.shard_route_unique(ids.reshape(-1), np.tile(vbase, b), W, lbits, sent.reshape(-1), np.tile(slot_rot, b))
    assert r.counts.tolist() == counts.tolist() and r.send_keys.tolist() == send.tolist()
    assert np.array_equal(r.pos.numpy(), O.shard_positions_from(pos, b, lens))
    blocks = field_positions(r.pos, b, lens)
    assert [tuple(t.shape) for t in blocks] == [(b,), (b, 4), (b,), (b, 6)]
    assert torch.equal(blocks[1] == 0, torch.from_numpy(cols[1] == 0))      # exactly the pads are unsent
    assert bool((blocks[0] > 0).all())                                       # SPARSE id 0 still travels


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepfm_b200.sharded import TorchDistComm
    comm = TorchDistComm()
    gen = torch.Generator().manual_seed(100 + rank)
    vocab = [13, 40, 9]
    lbits = 8
    ids = torch.stack([torch.randint(0, v, (20,), generator=gen) for v in vocab], dim=1)
    lv = [local_rows(v, world) for v in vocab]
    vbase = torch.tensor([0, lv[0], lv[0] + lv[1]])
    r = route_unique(ids, vbase, world, lbits=lbits)
    send_counts, recv_counts = comm.exchange_counts(r.counts)
    assert send_counts == r.counts.tolist()
    recv_keys = comm.all_to_all(r.send_keys, send_counts, recv_counts)
    # owner check: a received key names a row of a table this rank owns: row < ceil(V / W) of its field
    k = recv_keys.long()
    field = (k[:, None] >= vbase[None, :]).sum(1) - 1
    ok_owner = bool(torch.all(k - vbase[field] < torch.tensor(lv)[field]))
    # reply with a function of (rank, key); the sender must get it back in send order
    reply = comm.all_to_all((k * 3 + 1 + 1000 * rank).float()[:, None].repeat(1, 2), recv_counts, send_counts)
    owner_of_send = torch.repeat_interleave(torch.arange(world), torch.tensor(send_counts))
    ok_reply = bool(torch.equal(reply[:, 0], r.send_keys.float() * 3 + 1 + 1000 * owner_of_send))
    # every slot reads ITS row out of the reply buffer through its position
    back = reply[:, 0][r.pos.view(3, 20).t().reshape(-1) - 1]     # slot order
    want = ((vbase[None, :] + ids // world) * 3 + 1 + 1000 * (ids % world)).reshape(-1).float()
    ok_slot = bool(torch.equal(back, want))
    out.put((rank, ok_owner, ok_reply, ok_slot, sum(recv_counts), int(r.counts.sum())))
    dist.destroy_process_group()


def test_all_to_all_plumbing_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] and r[3] for r in res), res
    assert sum(r[4] for r in res) == sum(r[5] for r in res) <= 2 * 20 * 3     # every unique key arrived somewhere, once


def test_sharded_module_bookkeeping_cpu():
    """Host logic of ShardedFeatureEmbedding (no kernels): which tables are sharded / replicated, local table
    sizes under the rotated owner rule, load_from_full, and the torch routing in the plan's position layout."""
    from deepfm_b200.sharded import ShardedFeatureEmbedding
    schema = _schema(8, multihot=True)
    W = 3
    names = list(schema.fields)
    mods = [ShardedFeatureEmbedding(schema, 8, W, r, replicate_below=60) for r in range(W)]
    m = mods[1]
    small = {n for n, f in schema.fields.items() if f.feature_type != FeatureType.DENSE and f.vocabulary_size <= 60}
    assert {names[i] for i in m._repl_idx} == small and small            # s0 (50), s2 (7), s4 (11), q0 (40), q2 (9)
    for i in m._repl_idx:                                                 # replicated: full table on every rank
        assert m.second_order_embeddings[names[i]].weight.shape[0] == schema.fields[names[i]].vocabulary_size
    for i in m._table_idx:                                                # sharded: the ranks' shards partition the ids
        V = schema.fields[names[i]].vocabulary_size
        # every shard is sized ceil(V / W) (a row's local key is then the same number on every rank) ...
        assert all(mm.second_order_embeddings[names[i]].weight.shape[0] == local_rows(V, W) for mm in mods)
        assert sum(owned_rows(V, W, r, field=i) for r in range(W)) == V      # ... and the owned ids partition the table
    assert [p.shape[0] > 0 for p in m.table_parameters()] and len(m.table_parameters()) == 2 * len(m._table_idx)
    # routing restatement: replicated fields send nothing, sharded fields follow (id + field) mod W
    rng = np.random.default_rng(3)
    b = 29
    ins = []
    for n, f in schema.fields.items():
        if f.feature_type == FeatureType.DENSE:
            ins.append(torch.zeros(b))
        elif f.feature_type == FeatureType.SEQUENCE:
            ins.append(torch.from_numpy(rng.integers(0, f.vocabulary_size, (b, f.max_length))))
        else:
            ins.append(torch.from_numpy(rng.integers(0, f.vocabulary_size, b)))
    r = m.route_torch(ins)
    assert r.pos.numel() == b * m._all_S
    for i, blk in zip(m._all_tables, field_positions(r.pos, b, m._all_lens)):
        if i in m._repl_idx:
            assert int(blk.abs().sum()) == 0
        else:
            sent = ins[i].reshape(blk.shape) != 0 if schema.fields[names[i]].feature_type == FeatureType.SEQUENCE \
                else torch.ones_like(blk, dtype=torch.bool)
            assert torch.equal(blk > 0, sent)
            owner = (ins[i].reshape(blk.shape) + i) % W
            bounds = torch.cumsum(r.counts, 0)
            got_owner = torch.bucketize(blk - 1, bounds, right=True)
            assert torch.equal(got_owner[sent], owner[sent])
            key = m._vbase[i] + ins[i].reshape(blk.shape) // W                # the slot's position names ITS row
            assert torch.equal(r.send_keys.long()[(blk - 1).clamp_min(0)][sent], key[sent])


def test_peer_exchange_layout_and_capacity_logic():
    """Host arithmetic of PeerExchange (no GPU): buffer regions, and the collective-safe capacity decision."""
    from deepfm_b200.sharded import PeerExchange
    px = PeerExchange.__new__(PeerExchange)                 # the constructor needs symmetric memory; the logic does not
    px.world, px.rank, px.dim, px.cap = 3, 1, 64, 1024
    px.peer_base = [1 << 30, 2 << 30, 3 << 30]
    stride = 1024 * 68 * 4                                   # one (kind, parity) block: (cap, 64) vectors + (cap, 4) scalars
    assert px.region(0, 0) == (2 << 30, (2 << 30) + 1024 * 64 * 4)
    assert px.region(0, 1) == ((2 << 30) + stride, (2 << 30) + stride + 1024 * 64 * 4)
    assert px.region(1, 0, rank=2)[0] == (3 << 30) + 2 * stride and px.region(1, 1, rank=0)[0] == (1 << 30) + 3 * stride
    assert all(a % 256 == 0 for a in px.region(1, 1, rank=0))    # 256-byte rows start on 256-byte boundaries
    ok = [[100, 200, 300], [50, 60, 70], [400, 100, 100]]   # matrix[s][r]: rows s sends to r
    assert px.fits(ok)                                       # max sent 600 (+ the zero row), max received 550
    assert not px.fits([[524, 499, 1], [0, 0, 0], [0, 0, 0]])   # 1024 sent + the zero row > capacity
    assert not px.fits([[400, 0, 0], [400, 0, 0], [400, 0, 0]])  # rank 0 would receive 1200 rows


def test_peer_memory_shim_matches_installed_torch():
    """deepfm_b200/_peer.py is the only user of torch's private symmetric-memory module: the names it relies on exist."""
    from deepfm_b200 import _peer
    import importlib
    if not _peer.available():
        pytest.skip("this torch build has no torch.distributed._symmetric_memory")
    symm = importlib.import_module("torch.distributed._symmetric_memory")
    assert callable(symm.empty) and callable(symm.rendezvous)


def test_sharding_needs_a_table_to_shard():
    from deepfm_b200.sharded import ShardedFeatureEmbedding
    with pytest.raises(NotImplementedError, match="nothing to shard"):
        ShardedFeatureEmbedding(_schema(8), 8, 2, 0, replicate_below=10 ** 6)


def _ckpt_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepfm_b200.layers.embedding import FeatureEmbedding
    from deepfm_b200.sharded import ShardedFeatureEmbedding
    torch.manual_seed(0)                                   # the same "checkpoint" on every rank
    schema = _schema(8, multihot=True)
    full = FeatureEmbedding(schema, 8)
    with torch.no_grad():
        for p in full.parameters():
            p.add_(0.1 * torch.randn_like(p))
    shard = ShardedFeatureEmbedding(schema, 8, world, rank, replicate_below=60)
    shard.load_from_full(full)
    sd = shard.full_state_dict()                           # collective: shards -> reference layout
    ref = full.state_dict()
    ok = set(sd) == set(ref) and all(sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]) for k in ref)
    again = FeatureEmbedding(schema, 8)
    again.load_state_dict(sd)                              # the unsharded module (reference keys) loads it
    out.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_checkpoint_round_trip_gloo_world2():
    """load_from_full -> full_state_dict is the identity on the reference-layout state_dict."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_ckpt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def _reducer_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepfm_b200.sharded import DenseGradReducer
    torch.manual_seed(0)                                   # identical replicas
    early = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    late = torch.nn.Linear(3, 6)                           # plays the embedding's DENSE-field Linears
    red = DenseGradReducer(list(early.parameters()), list(late.parameters()), world)
    ok = True
    for step in range(2):                                  # twice: the hook counters must reset
        gen = torch.Generator().manual_seed(10 * step + rank)
        x = torch.randn(7, 3, generator=gen)
        for p in list(early.parameters()) + list(late.parameters()):
            p.grad = None
        early(late(x)).sum().backward()
        mine = [p.grad.clone() for p in list(early.parameters()) + list(late.parameters())]
        red.finish()
        for p, g in zip(list(early.parameters()) + list(late.parameters()), mine):
            parts = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(parts, g)
            ok = ok and torch.allclose(p.grad, sum(parts) / world, atol=1e-6)
    out.put((rank, ok))
    dist.destroy_process_group()


def _reducer_callback_worker(rank, world, port, out):
    """The product wiring: no per-parameter hooks -- the embedding's backward calls ``launch`` when it starts, and the
    embedding's own data-parallel gradients are views of ONE flat buffer that is all-reduced in place."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepfm_b200.sharded import DenseGradReducer

    class _Emb:                                           # what DenseGradReducer needs from ShardedFeatureEmbedding
        forward_fused = True
        on_backward_start = None
        dense_grad_flat = None

    torch.manual_seed(0)
    early = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    late = torch.nn.Linear(3, 6)
    emb = _Emb()
    red = DenseGradReducer(list(early.parameters()), list(late.parameters()), world, embedding=emb)
    ok = emb.on_backward_start is not None and not any(p._post_accumulate_grad_hooks for p in early.parameters()
                                                       if getattr(p, "_post_accumulate_grad_hooks", None))
    for step in range(2):
        gen = torch.Generator().manual_seed(10 * step + rank)
        x = torch.randn(7, 3, generator=gen)
        for p in list(early.parameters()) + list(late.parameters()):
            p.grad = None
        h = late(x)
        h.register_hook(lambda g: (emb.on_backward_start(), g)[1])      # fires where the embedding's backward node would start
        early(h).sum().backward()
        mine = [p.grad.clone() for p in list(early.parameters()) + list(late.parameters())]
        ok = ok and red._work is not None                                 # launched from the callback, before finish()
        # the embedding's gradients live in one flat bucket (views), like ShardedFeatureEmbedding.local_grads makes them
        lp = list(late.parameters())
        flat = torch.cat([p.grad.reshape(-1) for p in lp])
        off = 0
        for p in lp:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        emb.dense_grad_flat = flat
        red.finish()
        ok = ok and all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in lp)   # reduced in place
        for p, g in zip(list(early.parameters()) + lp, mine):
            parts = [torch.empty_like(g) for _ in range(world)]
            dist.all_gather(parts, g)
            ok = ok and torch.allclose(p.grad, sum(parts) / world, atol=1e-6)
    out.put((rank, ok))
    dist.destroy_process_group()


def test_dense_grad_reducer_callback_and_flat_bucket_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35000 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_callback_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def test_dense_grad_reducer_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


# ------------------------------------------------------------------------------------------ GPU

def _schema(D, multihot=False):
    fields = {}
    for i in range(3):
        fields[f"d{i}"] = FieldSchema(f"d{i}", FeatureType.DENSE, embedding_dim=D)
    for i, v in enumerate((50, 2000, 7, 301, 11)):
        fields[f"s{i}"] = FieldSchema(f"s{i}", FeatureType.SPARSE, vocabulary_size=v, embedding_dim=D)
    if multihot:
        for i, (v, L, comb) in enumerate(((40, 5, "mean"), (3000, 16, "sum"), (9, 3, "mean"))):
            fields[f"q{i}"] = FieldSchema(f"q{i}", FeatureType.SEQUENCE, vocabulary_size=v, embedding_dim=D,
                                          max_length=L, combiner=comb)
    return DatasetSchema(fields=fields)


def _exchange(bufs, counts, reverse=False):
    """bufs[s]: send buffer of rank s; counts[s][r]: items s sends to r (forward direction).
    reverse=True routes owner replies back (bufs[r] is ordered by source s)."""
    W = len(bufs)
    out = []
    if not reverse:
        offs = [np.concatenate([[0], np.cumsum(c)]) for c in counts]
        for r in range(W):
            out.append(torch.cat([bufs[s][offs[s][r]:offs[s][r + 1]] for s in range(W)]))
    else:
        roffs = [np.concatenate([[0], np.cumsum([counts[s][r] for s in range(W)])]) for r in range(W)]
        for s in range(W):
            out.append(torch.cat([bufs[r][roffs[r][s]:roffs[r][s + 1]] for r in range(W)]))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("W,D,multihot,repl", [(2, 32, False, 0), (3, 64, False, 0), (2, 64, True, 0), (4, 32, True, 0),
                                                 (3, 64, False, 60), (2, 32, True, 60)])
def test_sharded_equals_unsharded_emulated_ranks(W, D, multihot, repl):
    """W ranks emulated in one process (the exchange is done by slicing staging buffers): the routing kernels against
    the torch restatement bit for bit, forward views bit-identical to the unsharded module on the concatenated batch,
    every gradient within 2e-5.  repl: tables of at most that many rows are replicated."""
    from deepfm_b200.layers.embedding import FeatureEmbedding
    from deepfm_b200.layers.fm import FMInteraction
    from deepfm_b200.layers.l2 import l2_penalty
    from deepfm_b200.sharded import ShardedFeatureEmbedding
    torch.manual_seed(0)
    rng = np.random.default_rng(W * D)
    schema = _schema(D, multihot)
    b, lam = 301, 1e-3
    full = FeatureEmbedding(schema, D).cuda()
    with torch.no_grad():
        for p in full.parameters():
            p.add_(0.05 * torch.randn_like(p))
    batches = []
    for r in range(W):
        bt = {}
        for n, f in schema.fields.items():
            if f.feature_type == FeatureType.DENSE:
                bt[n] = torch.from_numpy(rng.uniform(-1, 1, b).astype(np.float32)).cuda()
            elif f.feature_type == FeatureType.SEQUENCE:
                x = (rng.zipf(1.3, (b, f.max_length)) - 1) % f.vocabulary_size
                x = x * (rng.random((b, f.max_length)) < 0.6)      # pads anywhere in the bag
                x[5] = 0                                           # an all-pad bag
                bt[n] = torch.from_numpy(x.astype(np.int64)).cuda()
            else:
                bt[n] = torch.from_numpy(((rng.zipf(1.3, b) - 1) % f.vocabulary_size).astype(np.int64)).cuda()
        batches.append(bt)
    whole = {n: torch.cat([bt[n] for bt in batches]) for n in schema.fields}
    T = schema.total_embedding_dim
    g_first = torch.randn(W * b, 1, device="cuda")
    g_flat = torch.randn(W * b, T, device="cuda")
    g_fm = torch.randn(W * b, 1, device="cuda") * 0.1
    fo, fe, fl = full(whole)
    fm = FMInteraction()(fe)
    ((fo * g_first).sum() + (fl * g_flat).sum() + (fm * g_fm).sum() + l2_penalty(full, lam)).backward()

    mods = [ShardedFeatureEmbedding(schema, D, W, r, replicate_below=repl, grad_scale=1.0).cuda() for r in range(W)]   # sum semantics
    for m in mods:
        m.load_from_full(full)
        m.grad_mode = "dense"
    ins = [m._prepare(bt) for m, bt in zip(mods, batches)]
    routes = [m.route(x) for m, x in zip(mods, ins)]
    counts = [r.counts.tolist() for r in routes]
    for m, x, r in zip(mods, ins, routes):          # routing kernels == the torch restatement, bit-exactly
        ref = m.route_torch(x)
        n_u = int(ref.counts.sum())
        assert torch.equal(r.counts, ref.counts)
        assert torch.equal(r.send_keys[:n_u], ref.send_keys)
        assert torch.equal(r.pos, ref.pos)
        assert n_u < b * m._S                         # Zipf ids: repeats travel once
        r.send_keys = r.send_keys[:n_u]
    recv_keys = _exchange([r.send_keys for r in routes], counts)
    gathered = [m.gather(k) for m, k in zip(mods, recv_keys)]              # (vec, sc, bkeys) staging
    got_v = _exchange([g[0] for g in gathered], counts, reverse=True)
    got_s = _exchange([g[1] for g in gathered], counts, reverse=True)
    got = []
    for m, v, c in zip(mods, got_v, got_s):         # row 0 of the reply buffers is the reserved zero row
        bv, bs = m.reply_buffers(v.shape[0], v)
        bv[1:] = v
        bs[1:] = c
        got.append((bv, bs))
    outs = [m.finish(x, r.pos, g[0], g[1], True) for m, x, r, g in zip(mods, ins, routes, got)]
    assert torch.equal(torch.cat([o[0] for o in outs]), fo.detach())      # same rows, same order: bit-identical
    assert torch.equal(torch.cat([o[2] for o in outs]), fl.detach())
    assert torch.equal(torch.cat([o[3] for o in outs]), fm.detach())
    # backward
    gscale = torch.ones((), device="cuda")
    staged, dense = [], []
    for r, (m, o) in enumerate(zip(mods, outs)):
        sl = slice(r * b, (r + 1) * b)
        params = m._ordered_params()
        gf, gl, gm = g_first[sl].contiguous(), g_flat[sl].contiguous(), g_fm[sl].contiguous()
        sv, ss = m.reduce_grads(routes[r], gf, None, gl, gm, o[1], o[4], o[5])
        n_u = routes[r].send_keys.numel()
        staged.append((sv[:n_u], ss[:n_u]))
        dense.append(m.local_grads(o[6], got[r][0], got[r][1], gf, None, gl, gm, o[1], o[2], o[4], params, lam / W, gscale,
                                   o[5], keys=o[7]))
    g_vec = _exchange([p[0] for p in staged], counts)
    g_sc = _exchange([p[1] for p in staged], counts)
    for r, m in enumerate(mods):
        params = m._ordered_params()
        bs = m.sort_owner_keys(gathered[r][2])
        tg = m.owner_backward(bs, g_vec[r], g_sc[r], params, lam, gscale)
        for i, g in tg.items():
            slot = m._slot_of_param[i]
            name = m.field_names[slot // 5]
            assert slot // 5 not in m._repl_idx
            ref = (full.second_order_embeddings if slot % 5 == 0 else full.first_order_embeddings)[name].weight.grad
            ref = ref[(r - slot // 5) % W::W]
            assert_close_rel(g[: ref.shape[0]].cpu(), ref.cpu(), 2e-5, f"rank {r} {name} slot {slot % 5}")
    # replicated parameters (DENSE-field Linears, small tables): the sum over ranks of the per-rank gradients is the
    # full gradient (each rank was given l2 / W)
    for i, tab in enumerate(mods[0]._param_is_table):
        if tab:
            continue
        slot = mods[0]._slot_of_param[i]
        name = mods[0].field_names[slot // 5]
        mod = (full.second_order_embeddings if slot % 5 < 2 else full.first_order_embeddings)[name]
        ref = (mod.weight if slot % 5 in (0, 2) else mod.bias).grad
        got_g = sum(p[i] for p in dense)
        assert_close_rel(got_g.cpu(), ref.cpu(), 2e-5, f"dense {name} slot {slot % 5}")
