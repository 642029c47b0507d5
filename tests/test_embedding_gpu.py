"""GPU parity of K1/K2 (FeatureEmbedding + FM + L2) through the C ABI against the golden fixtures of
the unmodified reference and against the numpy oracle.  fp32 tolerance: per-tensor max-norm relative
error <= 1e-5 forward, <= 1e-4 gradients; integer artefacts bit-exact."""

import ctypes as C

import numpy as np
import pytest
import torch

from deepfm_b200 import _lib
from deepfm_b200.layers.embedding import FeatureEmbedding
from deepfm_b200.layers.fm import FMInteraction
from deepfm_b200.layers.l2 import l2_penalty
from deepfm_b200.schema import DatasetSchema, FeatureType, FieldSchema
from oracle import deepfm_oracle as O
from tests.golden import spec
from tests.helpers import assert_close_rel, grads_of, load_golden, load_params, split_prefixed, to_dev

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-5, 1e-4


def golden_embedding():
    g = load_golden("embedding.npz")
    emb = FeatureEmbedding(spec.golden_schema(), fm_embed_dim=spec.FM_DIM)
    load_params(emb, split_prefixed(g, "param/"))
    return g, emb.cuda()


def test_forward_three_views_match_reference():
    g, emb = golden_embedding()
    fo, fe, fl = emb(to_dev(split_prefixed(g, "batch/")))
    assert fo.shape == (6, 1) and fe.shape == (6, 7, spec.FM_DIM) and fl.shape == (6, 40)
    assert fo.is_contiguous() and fe.is_contiguous() and fl.is_contiguous()
    assert_close_rel(fo.detach().cpu(), g["first_order"], FWD_TOL, "first_order")
    assert_close_rel(fe.detach().cpu(), g["field_embeddings"], FWD_TOL, "field_embeddings")
    assert_close_rel(fl.detach().cpu(), g["flat"], FWD_TOL, "flat")
    # a plain gather of stored rows is exact: the un-projected sparse column is bit-identical
    assert np.array_equal(fl.detach().cpu().numpy()[:, :8], g["flat"][:, :8])


def test_backward_matches_reference_autograd_including_l2():
    g, emb = golden_embedding()
    fo, fe, fl = emb(to_dev(split_prefixed(g, "batch/")))
    up = to_dev({k: g[k] for k in ("g_first", "g_field", "g_flat")})
    loss = (fo * up["g_first"]).sum() + (fe * up["g_field"]).sum() + (fl * up["g_flat"]).sum() \
        + l2_penalty(emb, spec.L2_REG)
    loss.backward()
    ref = split_prefixed(g, "grad/")
    got = grads_of(emb)
    assert set(got) == set(ref)
    for k in ref:
        assert got[k] is not None, k
        assert_close_rel(got[k], ref[k], GRAD_TOL, k)
    # padding row: no lookup gradient, only the L2 term (SURVEY a3')
    np.testing.assert_allclose(got["second_order_embeddings.u.weight"][0], 2 * spec.L2_REG * 0.25, rtol=1e-6)


def test_l2_value_and_standalone_gradient():
    g, emb = golden_embedding()
    val = l2_penalty(emb, spec.L2_REG)
    assert abs(val.item() - float(g["l2_loss"])) < 1e-5 * float(g["l2_loss"])
    val.backward()                       # no embedding forward in the graph: direct 2*lambda*p
    for k, p in emb.named_parameters():
        assert_close_rel(p.grad.cpu(), 2 * spec.L2_REG * p.detach().cpu().numpy(), 1e-6, k)


def test_backward_without_l2_and_with_missing_upstreams():
    g, emb = golden_embedding()
    params, batch = split_prefixed(g, "param/"), split_prefixed(g, "batch/")
    fo, fe, fl = emb(to_dev(batch))
    gfl = torch.from_numpy(g["g_flat"]).cuda()
    (fl * gfl).sum().backward()          # only the flat view is used: g_first, g_field are None
    ref = O.embedding_backward(spec.golden_schema(), params, batch, spec.FM_DIM, np.zeros_like(g["g_first"]),
                               np.zeros_like(g["g_field"]), g["g_flat"], l2_reg=0.0)
    got = grads_of(emb)
    for k in ref:
        assert_close_rel(got[k], ref[k], GRAD_TOL, k)


def test_fused_fm_value_and_gradient_match_oracle():
    g, emb = golden_embedding()
    params, batch = split_prefixed(g, "param/"), split_prefixed(g, "batch/")
    fo, fe, fl = emb(to_dev(batch))
    fm = FMInteraction()(fe)             # served by K1 (attribute on the tensor), not recomputed
    assert fm is fe._dfm_fm[0]
    assert_close_rel(fm.cpu().detach(), O.fm_forward(g["field_embeddings"]), FWD_TOL, "fm")
    gfm = np.linspace(-1, 1, 6, dtype=np.float32)[:, None]
    (fm * torch.from_numpy(gfm).cuda()).sum().backward()
    ge = O.fm_backward(g["field_embeddings"], gfm)
    ref = O.embedding_backward(spec.golden_schema(), params, batch, spec.FM_DIM, np.zeros_like(g["g_first"]),
                               ge, np.zeros_like(g["g_flat"]), l2_reg=0.0)
    got = grads_of(emb)
    for k in ref:
        assert_close_rel(got[k], ref[k], GRAD_TOL, k)


def test_all_zero_indices_give_exactly_zero_views():
    # reference tests/test_layers.py:43-51 (sparse-only schema of tests/test_layers.py:13-27)
    fields = {n: FieldSchema(n, FeatureType.SPARSE, vocabulary_size=v, embedding_dim=d)
              for n, v, d in (("user_id", 100, 8), ("item_id", 200, 16), ("genre", 20, 4))}
    emb = FeatureEmbedding(DatasetSchema(fields=fields), fm_embed_dim=16).cuda()
    batch = {n: torch.zeros(4, dtype=torch.long, device="cuda") for n in fields}
    fo, fe, fl = emb(batch)
    assert fo.shape == (4, 1) and fe.shape == (4, 3, 16) and fl.shape == (4, 28)
    assert fo.abs().sum().item() == 0 and fe.abs().sum().item() == 0 and fl.abs().sum().item() == 0
    batch = {"user_id": torch.tensor([1, 5, 10, 50]), "item_id": torch.tensor([2, 10, 100, 150]),
             "genre": torch.tensor([1, 3, 7, 15])}
    fo, fe, fl = emb({k: v.cuda() for k, v in batch.items()})
    (fo.sum() + fe.sum() + fl.sum()).backward()
    for name, p in emb.named_parameters():       # tests/test_layers.py:53-62
        assert p.grad is not None, name


def test_empty_batch():
    g, emb = golden_embedding()
    batch = {k: v[:0] for k, v in to_dev(split_prefixed(g, "batch/")).items()}
    fo, fe, fl = emb(batch)
    assert fo.shape == (0, 1) and fe.shape == (0, 7, spec.FM_DIM) and fl.shape == (0, 40)
    (fo.sum() + fe.sum() + fl.sum()).backward()
    for k, p in emb.named_parameters():
        assert p.grad is not None and p.grad.abs().sum().item() == 0, k


def test_out_of_range_index_raises_like_reference():
    g, emb = golden_embedding()
    emb.check_indices = "sync"            # check inside forward (one host sync per call)
    batch = to_dev(split_prefixed(g, "batch/"))
    batch["u"] = batch["u"].clone()
    batch["u"][2] = 11                    # vocabulary_size == 11
    with pytest.raises(IndexError):
        emb(batch)


def test_out_of_range_index_raises_lazily_by_default():
    """Default: no host sync on the hot path; the status word travels asynchronously and the error surfaces at
    raise_if_bad_index() / a later forward (never silently)."""
    g, emb = golden_embedding()
    assert emb.check_indices is True
    good = to_dev(split_prefixed(g, "batch/"))
    bad = dict(good)
    bad["u"] = good["u"].clone()
    bad["u"][2] = -3
    emb(bad)                              # clamped to the padding row, flagged
    with pytest.raises(IndexError):
        emb.raise_if_bad_index()
    emb(good)
    emb.raise_if_bad_index()              # the flag was cleared, a clean batch stays clean
    emb(bad)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        emb(good)                         # surfaces at the next forward at the latest once it has arrived


# ----------------------------------------------------------------- integer artefacts (bit-exact)

def _keys_on_device(emb, batch_np):
    lib = _lib.lib()
    plan = emb._ensure_plan()
    names = emb.field_names
    ins = [torch.from_numpy(batch_np[n]).cuda() for n in names]
    B = len(batch_np[names[0]])
    keys = torch.empty(B * emb._S, dtype=torch.int32, device="cuda")
    _lib.check(lib.dfm_emit_keys(plan, B, _lib.ptr_array(ins), keys.data_ptr(), _lib.stream_ptr()))
    ws = torch.empty(lib.dfm_embed_bwd_workspace_bytes(plan, B) + 1024, dtype=torch.uint8, device="cuda")
    sk, sp = torch.empty_like(keys), torch.empty_like(keys)
    _lib.check(lib.dfm_sort_keys(plan, keys.numel(), keys.data_ptr(), sk.data_ptr(), sp.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    u32 = lambda t: t.cpu().numpy().view(np.uint32)
    return u32(keys), u32(sk), u32(sp)


def test_keys_sort_and_segments_bit_exact_vs_oracle():
    schema = spec.golden_schema()
    emb = FeatureEmbedding(schema, fm_embed_dim=spec.FM_DIM).cuda()
    rng = np.random.default_rng(5)
    B = 3000
    batch = {}
    for n, f in schema.fields.items():
        k = O._kind(f)
        if k == "dense":
            batch[n] = rng.random(B).astype(np.float32)
        elif k == "sparse":
            batch[n] = rng.integers(0, f.vocabulary_size, B).astype(np.int64)
        else:
            ids = rng.integers(0, f.vocabulary_size, (B, f.max_length)).astype(np.int64)
            ids[rng.random((B, f.max_length)) < 0.4] = 0
            batch[n] = ids
    keys, sk, sp = _keys_on_device(emb, batch)
    ok, opay = O.emit_keys(schema, batch)
    pad = np.uint32(O.slot_layout(schema)[2][-1])
    ok = np.where(ok == O.PAD_KEY, pad, ok)        # device PAD key = total rows (fewer sort bits)
    assert np.array_equal(keys, ok)
    rk, rp = O.sort_pairs(ok, opay)
    assert np.array_equal(sk, rk) and np.array_equal(sp, rp)
    # segment counts through the backward's device counters
    inputs = {n: torch.from_numpy(v).cuda() for n, v in batch.items()}
    fo, fe, fl = emb(inputs)
    (fo.sum() + fl.sum()).backward()
    n_valid, n_unique = emb.last_counts.cpu().tolist()
    uk, starts = O.segment_heads(np.where(rk == pad, O.PAD_KEY, rk))
    assert n_valid == int(starts[-1]) and n_unique == len(uk)


# ----------------------------------------------------------------- larger shapes vs the oracle

def _criteo_like(n_sparse=5, n_dense=3, D=16, vocab=(50, 2000, 7, 300, 11), seqs=()):
    fields = {}
    for i in range(n_dense):
        fields[f"d{i}"] = FieldSchema(f"d{i}", FeatureType.DENSE, embedding_dim=D)
    for i in range(n_sparse):
        fields[f"s{i}"] = FieldSchema(f"s{i}", FeatureType.SPARSE, vocabulary_size=vocab[i % len(vocab)], embedding_dim=D)
    for i, (L, comb) in enumerate(seqs):
        fields[f"q{i}"] = FieldSchema(f"q{i}", FeatureType.SEQUENCE, vocabulary_size=40, embedding_dim=D,
                                      max_length=L, combiner=comb)
    return DatasetSchema(fields=fields)


def _random_batch(schema, B, rng, zipf=True):
    batch = {}
    for n, f in schema.fields.items():
        k = O._kind(f)
        if k == "dense":
            batch[n] = rng.uniform(-1, 1, B).astype(np.float32)
        else:
            shape = (B,) if k == "sparse" else (B, f.max_length)
            ids = (rng.zipf(1.3, shape) - 1) % f.vocabulary_size if zipf else rng.integers(0, f.vocabulary_size, shape)
            batch[n] = ids.astype(np.int64)
    return batch


@pytest.mark.parametrize("D,B,seqs", [(16, 777, ()), (64, 2048, ()), (8, 513, ((6, "mean"), (4, "sum"), (3, "max")))])
def test_aliased_vector_path_vs_oracle_with_long_segments(D, B, seqs):
    """All dims == D (field_embeddings aliases flat), Zipf ids => hot rows whose segments span many
    32-position chunks (exercises the stitch pass); compares fwd, FM and dense grads with the oracle."""
    torch.manual_seed(0)
    rng = np.random.default_rng(D + B)
    schema = _criteo_like(D=D, seqs=seqs)
    emb = FeatureEmbedding(schema, fm_embed_dim=D).cuda()
    with torch.no_grad():
        for p in emb.parameters():
            p.add_(0.05 * torch.randn_like(p))       # non-zero row 0 and biases
    params = {k: v.detach().cpu().numpy() for k, v in emb.state_dict().items()}
    batch = _random_batch(schema, B, rng)
    fo, fe, fl = emb(to_dev(batch))
    if not seqs:
        assert fe.data_ptr() == fl.data_ptr()            # written once
    ref = O.embedding_forward(schema, params, batch, D)
    assert_close_rel(fo.detach().cpu(), ref["first_order"], FWD_TOL, "first_order")
    assert_close_rel(fe.detach().cpu(), ref["field_embeddings"], FWD_TOL, "field_embeddings")
    assert_close_rel(fl.detach().cpu(), ref["flat"], FWD_TOL, "flat")
    fm = FMInteraction()(fe)
    assert_close_rel(fm.detach().cpu(), O.fm_forward(ref["field_embeddings"].astype(np.float64)), 2e-5, "fm")
    g_first = rng.standard_normal((B, 1)).astype(np.float32)
    g_flat = rng.standard_normal(fl.shape).astype(np.float32)
    g_fm = rng.standard_normal((B, 1)).astype(np.float32) * 0.1
    loss = (fo * torch.from_numpy(g_first).cuda()).sum() + (fl * torch.from_numpy(g_flat).cuda()).sum() \
        + (fm * torch.from_numpy(g_fm).cuda()).sum() + l2_penalty(emb, 1e-3)
    loss.backward()
    ge = O.fm_backward(ref["field_embeddings"], g_fm)
    want = O.embedding_backward(schema, params, batch, D, g_first, ge, g_flat, l2_reg=1e-3)
    got = grads_of(emb)
    for k in want:
        assert_close_rel(got[k], want[k], GRAD_TOL, k)


def test_row_sparse_mode_equals_dense_on_touched_rows_and_is_deterministic():
    torch.manual_seed(1)
    rng = np.random.default_rng(11)
    D, B = 32, 4096
    schema = _criteo_like(D=D, n_sparse=6, seqs=((5, "mean"),))
    emb = FeatureEmbedding(schema, fm_embed_dim=D).cuda()
    batch = to_dev(_random_batch(schema, B, rng))
    g_flat = torch.randn(B, schema.total_embedding_dim, device="cuda")

    def run(mode):
        emb.grad_mode = mode
        emb.zero_grad(set_to_none=True)
        fo, fe, fl = emb(batch)
        fm = FMInteraction()(fe)
        ((fl * g_flat).sum() + fo.sum() + fm.sum() * 0.01 + l2_penalty(emb, 1e-3)).backward()
        return emb

    run("dense")
    dense = {k: p.grad.clone() for k, p in emb.named_parameters()}
    run("row_sparse")
    rs = emb.row_grads
    assert rs is not None
    for name, (rows, g2, g1) in rs.per_table().items():
        d2 = dense[f"second_order_embeddings.{name}.weight"]
        d1 = dense[f"first_order_embeddings.{name}.weight"]
        assert torch.equal(d2[rows], g2) and torch.equal(d1[rows, 0], g1), name
        assert emb.second_order_embeddings[name].weight.grad is None      # tables get no dense grad
    for k, p in emb.named_parameters():                                    # non-table params stay dense
        if p.grad is not None:
            assert torch.equal(p.grad, dense[k]), k
    emb.materialize_sparse_grads()
    w = emb.second_order_embeddings["s1"].weight
    assert w.grad.is_sparse and w.grad.shape == w.shape
    # bit-reproducible: no float atomics anywhere in K2
    a = {k: v.clone() for k, v in dense.items()}
    run("dense")
    for k, p in emb.named_parameters():
        assert torch.equal(p.grad, a[k]), k


# ----------------------------------------------------------------- FM stand-alone kernels

def test_fm_golden_and_reference_test_facts():
    g = load_golden("fm.npz")
    fm = FMInteraction()
    assert sum(p.numel() for p in fm.parameters()) == 0                    # tests/test_layers.py:75-77
    ex = torch.tensor([[[1., 2.], [3., 4.], [5., 6.]]], device="cuda")
    assert fm(ex).item() == 67.0                                            # notes/deepfm.md:72-91
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    out = fm(x)
    assert out.shape == (5, 1)
    assert_close_rel(out.detach().cpu(), g["out"], FWD_TOL, "fm out")
    out.backward(torch.from_numpy(g["g"]).cuda())
    assert_close_rel(x.grad.cpu(), g["gx"], GRAD_TOL, "fm grad")
    x = torch.randn(2, 4, 8, device="cuda")                                 # tests/test_layers.py:79-92
    want = torch.zeros(2, 1, device="cuda")
    for i in range(4):
        for j in range(i + 1, 4):
            want += (x[:, i] * x[:, j]).sum(dim=1, keepdim=True)
    assert torch.allclose(fm(x), want, atol=1e-5, rtol=1e-5)
    assert torch.all(fm(torch.randn(2, 1, 8, device="cuda")) == 0)          # tests/test_layers.py:94-98
    assert fm(torch.randn(4, 3, 16, device="cuda")).shape == (4, 1)         # tests/test_layers.py:69-73
